// Device-side data layout shared by the kernels of the gen_data hot path.
//
// Device voxel index = the OUTPUT order of saveVoxelGrid (voxelize_utils.cpp:229-255):
//   L = (x * Sy + y) * Sz + z      (z fastest)
// so `.label` is the dense label array as is, the bit-packers write consecutive words, and with
// Sz = 32 one (x,y) column is one 32-bit word of every bitmap.  The reference's internal x-fastest
// index (VoxelGrid.h:117) never exists on the device; only its visit ORDER does (vx_visibility.cu).
#ifndef VOXELIZER_B200_VX_DEVICE_CUH_
#define VOXELIZER_B200_VX_DEVICE_CUH_

#include <cuda_runtime.h>
#include <stdint.h>

#include "vx_dda.h"

#define VX_MAX_PROBES 4096u
#define VX_SLOT_STATS 16

// Open-addressing histogram of one grid of one frame: (L << 16 | label16) -> count, and the dense
// per-voxel vote array best[L] = max over labels of (count << 16 | 0xFFFF - label).  Keys are stored + 1
// so that an all-zero table is empty; best[] is zero outside a vote (the winners reset their entries).
struct VxTable {
  unsigned long long* keys;   // [cap]  ((L << 16 | label) + 1), 0 = free
  uint32_t* counts;           // [cap]
  unsigned long long* best;   // [nvox]
  uint32_t* overflow;         // [1] set when a probe sequence ran out
  uint32_t log2_cap;
};

// One frame of a batch.  The histogram tables are shared by frames that are `table_sets` apart (they are
// processed by different fill -> vote -> emit launches); the rest is per frame.
struct VxSlot {
  VxTable input;              // priorGrid histogram
  VxTable target;             // pastGrid histogram
  uint32_t* occ;              // [words] target grid occupancy (count > 0), bit L
  uint32_t* occl;             // [words] occlusions_ > -1
  uint32_t* alive;            // [words] free cells with invalid_ != -1 (the traversal's live set)
  uint16_t* out_label;        // [nvox]
  uint32_t* out_bin;          // [words] MSB-first packed
  uint32_t* out_occluded;     // [words]
  uint32_t* out_invalid;      // [words]
  unsigned long long* counters;  // [VX_SLOT_STATS] rays, steps, waves, resolve waves, rounds, cycles[8]
  // insertOcclusionLabels (vx_options.occlusion_labels; null otherwise)
  uint32_t* occ_filled;       // [words] occupancy after the fill: what the invalid passes see
  uint32_t* occluder;         // [nvox] pass 0: reference-internal index of the voxel that occludes a free voxel (last visit)
  uint32_t* fill_source;      // [nvox] L of the voxel a filled voxel copied from, 0xFFFFFFFF otherwise
};

// Scratch of one RESIDENT visibility CTA (claimed at run time, so the pool is as large as the number
// of CTAs the device can hold, not as the number of frames of a batch).
struct VxVisScratch {
  uint32_t* memo;     // [nvox]  occludedBy_ as (pass epoch << 25 | id of the latest ray that crossed the cell)
  uint32_t* hitbits;  // [visits / 32 + 1] ray id -> the ray ended on an occupied cell (current pass)
  uint32_t* hitcell;  // [visits] ray id -> L of that cell (pass 0 of frames with insertOcclusionLabels; null without the option)
};

// Fill work is organised by SCAN: a tile of a scan is loaded once and inserted into every frame of the
// launch whose window holds that scan (a scan sits in ~14 windows at stride 5, 71 at stride 1).
struct VxScanWork {
  const float4* points;
  const uint32_t* labels;
  uint32_t n;
  uint32_t use_begin;  // into the launch's VxScanUse array
  uint32_t use_count;
  uint32_t pad;
};

struct VxScanUse {
  uint32_t slot;      // frame of the batch
  uint32_t ap_index;  // into the batch's ap array (16 floats each)
  uint32_t both;      // 1: prior scan -> input AND target grid; 0: past scan -> target only
};

struct VxFilterDev {
  float min_range, max_range;
  int32_t hidecar;
  int32_t has_label_filter;
  const uint32_t* reject_bits;  // 65536-bit LUT over the (masked) label: joined label is in `ignore`
};

// Per-frame visibility work.
struct VxVisFrame {
  uint32_t n_endpoints;
  uint32_t endpoint_offset;  // into the batch's endpoint array (3 floats each)
  uint32_t flags;            // vx_frame_flags
  uint32_t pad;
};

__device__ __forceinline__ uint32_t vx_ldcg_u32(const uint32_t* p) { return __ldcg(p); }

// Fire-and-forget reductions (SASS REDG).  Written in PTX because the compiler turns a result-less atomicXxx()
// inside divergent code into the returning form (ATOM ..., RZ), which the warp then waits for.
__device__ __forceinline__ void vx_red_add(uint32_t* p, uint32_t v) { asm volatile("red.relaxed.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v)); }
__device__ __forceinline__ void vx_red_max(uint32_t* p, uint32_t v) { asm volatile("red.relaxed.gpu.global.max.u32 [%0], %1;" ::"l"(p), "r"(v)); }
__device__ __forceinline__ void vx_red_or(uint32_t* p, uint32_t v) { asm volatile("red.relaxed.gpu.global.or.b32 [%0], %1;" ::"l"(p), "r"(v)); }
__device__ __forceinline__ void vx_red_and(uint32_t* p, uint32_t v) { asm volatile("red.relaxed.gpu.global.and.b32 [%0], %1;" ::"l"(p), "r"(v)); }
__device__ __forceinline__ void vx_red_max64(unsigned long long* p, unsigned long long v) {
  asm volatile("red.relaxed.gpu.global.max.u64 [%0], %1;" ::"l"(p), "l"(v));
}

#endif
