// K1 + K2 (fill), K3 (vote), K5 (emit / pack): the HBM-bound passes of the gen_data hot path.
//
//   fill  = fillVoxelGrid + VoxelGrid::insert     voxelize_utils.cpp:173-203, VoxelGrid.cpp:45-63
//   vote  = majority label of saveVoxelGrid        voxelize_utils.cpp:237-245
//   emit  = .label / .bin output + pack()          voxelize_utils.cpp:205-216, 257-294
//   pack  = .occluded / .invalid output            voxelize_utils.cpp:267-283, VoxelGrid.cpp:65-73
//
// All fp32 arithmetic that decides which voxel a point falls into is written with the
// round-to-nearest intrinsics (__fmul_rn, __fadd_rn, __fdiv_rn, __fsqrt_rn): the compiler never
// contracts those into FMAs, which is what reproduces the reference's SSE2 (FMA-free) Eigen code.
#include "vx_launch.h"

namespace {

constexpr int kFillThreads = 256;
constexpr int kFillPointsPerThread = 4;
constexpr int kFillTile = kFillThreads * kFillPointsPerThread;

// (voxel,label) -> count, open addressing with linear probing.  Keys go EMPTY -> key exactly once per
// frame, so a (possibly L1-stale) plain load is safe: a stale EMPTY is corrected by the CAS.
__device__ __forceinline__ void table_add(const VxTable& t, unsigned long long key, uint32_t n) {
  const uint32_t mask = (1u << t.log2_cap) - 1u;
  uint32_t h = (uint32_t)((key * 0x9E3779B97F4A7C15ull) >> (64u - t.log2_cap));
  for (uint32_t probe = 0; probe < VX_MAX_PROBES; ++probe) {
    unsigned long long k = t.keys[h];
    if (k == VX_EMPTY_KEY) {
      k = atomicCAS(&t.keys[h], VX_EMPTY_KEY, key);
      if (k == VX_EMPTY_KEY) {
        const uint32_t u = atomicAdd(t.used_count, 1u);
        t.used[u] = h;
        k = key;
      }
    }
    if (k == key) {
      atomicAdd(&t.counts[h], n);
      return;
    }
    h = (h + 1u) & mask;
  }
  atomicExch(t.overflow, 1u);
}

__global__ void __launch_bounds__(kFillThreads)
vx_fill_kernel(const VxFillItem* __restrict__ items, const float* __restrict__ ap_all, const VxSlot* __restrict__ slots,
               const VxGridGeom g, const VxFilterDev f) {
  const VxFillItem it = items[blockIdx.y];
  const uint32_t base = blockIdx.x * kFillTile;
  if (base >= it.n) return;

  const float* __restrict__ m = ap_all + 16u * it.ap_index;  // column-major: m[c*4 + r]
  const float m00 = m[0], m10 = m[1], m20 = m[2];
  const float m01 = m[4], m11 = m[5], m21 = m[6];
  const float m02 = m[8], m12 = m[9], m22 = m[10];
  const float m03 = m[12], m13 = m[13], m23 = m[14];
  const VxSlot& slot = slots[it.slot];

  float4 p[kFillPointsPerThread];
  uint32_t lab[kFillPointsPerThread];
  bool valid[kFillPointsPerThread];
#pragma unroll
  for (int u = 0; u < kFillPointsPerThread; ++u) {
    const uint32_t idx = base + u * kFillThreads + threadIdx.x;
    valid[u] = idx < it.n;
    if (valid[u]) {
      p[u] = __ldcs(it.points + idx);   // streamed once per frame: evict-first
      lab[u] = __ldcs(it.labels + idx);
    }
  }

#pragma unroll
  for (int u = 0; u < kFillPointsPerThread; ++u) {
    bool keep = valid[u];
    unsigned long long key = 0;
    if (keep) {
      const float x = p[u].x, y = p[u].y, z = p[u].z;
      // voxelize_utils.cpp:188  Eigen norm: sqrtf(x*x + (y*y + z*z))
      const float range = __fsqrt_rn(__fadd_rn(__fmul_rn(x, x), __fadd_rn(__fmul_rn(y, y), __fmul_rn(z, z))));
      if (range < f.min_range || range > f.max_range) keep = false;  // NaN range passes, as in the reference
      // voxelize_utils.cpp:190  ego-vehicle box
      if (f.hidecar && x < 3.0f && x > -2.0f && fabsf(y) < 2.0f) keep = false;
      if (keep && f.has_label_filter) {  // voxelize_utils.cpp:195-198 (join, then ignore) as a 64 Ki-bit LUT
        const uint32_t l16 = lab[u] & 0xFFFFu;
        if ((__ldg(f.reject_bits + (l16 >> 5)) >> (l16 & 31u)) & 1u) keep = false;
      }
      if (keep) {
        // voxelize_utils.cpp:193  ((m_r0*x + m_r1*y) + m_r2*z) + m_r3
        const float px = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(m00, x), __fmul_rn(m01, y)), __fmul_rn(m02, z)), m03);
        const float py = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(m10, x), __fmul_rn(m11, y)), __fmul_rn(m12, z)), m13);
        const float pz = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(m20, x), __fmul_rn(m21, y)), __fmul_rn(m22, z)), m23);
        // VoxelGrid.cpp:46-53  floor((p - offset) / resolution), IEEE division; NaN / out of range rejected
        const float fi = floorf(__fdiv_rn(__fsub_rn(px, g.off[0]), g.res));
        const float fj = floorf(__fdiv_rn(__fsub_rn(py, g.off[1]), g.res));
        const float fk = floorf(__fdiv_rn(__fsub_rn(pz, g.off[2]), g.res));
        keep = (fi >= 0.0f) && (fi < (float)g.nx) && (fj >= 0.0f) && (fj < (float)g.ny) && (fk >= 0.0f) && (fk < (float)g.nz);
        if (keep) {
          const uint32_t L = ((uint32_t)fi * (uint32_t)g.ny + (uint32_t)fj) * (uint32_t)g.nz + (uint32_t)fk;
          key = ((unsigned long long)L << 16) | (unsigned long long)(lab[u] & 0xFFFFu);
        }
      }
    }
    // K2: one atomic per distinct (voxel,label) per warp.  Scan order is beam-major / azimuth-minor,
    // so neighbouring lanes mostly share a voxel.
    const uint32_t active = __ballot_sync(0xFFFFFFFFu, keep);
    if (keep) {
      const uint32_t peers = __match_any_sync(active, key);
      if ((uint32_t)(__ffs(peers) - 1) == (threadIdx.x & 31u)) {
        const uint32_t n = (uint32_t)__popc(peers);
        table_add(slot.target, key, n);
        if (it.both) table_add(slot.input, key, n);
      }
    }
  }
}

// K3.  best[L] = max over labels of (count << 16 | 0xFFFF - label): largest count, ties -> smallest
// label -- the result of scanning the reference's std::map in ascending label order with a strict '>'.
__global__ void __launch_bounds__(256) vx_vote_kernel(const VxSlot* __restrict__ slots, int keep_tables) {
  const VxSlot& s = slots[blockIdx.y >> 1];
  const bool is_target = (blockIdx.y & 1u) != 0;
  const VxTable& t = is_target ? s.target : s.input;
  unsigned long long* best = is_target ? s.best_target : s.best_input;
  const uint32_t used = *t.used_count;
  for (uint32_t u = blockIdx.x * blockDim.x + threadIdx.x; u < used; u += gridDim.x * blockDim.x) {
    const uint32_t h = t.used[u];
    const unsigned long long key = t.keys[h];
    const unsigned long long c = t.counts[h];
    const unsigned long long L = key >> 16;
    const unsigned long long label = key & 0xFFFFull;
    atomicMax(&best[L], (c << 16) | (0xFFFFull - label));
    if (!keep_tables) {
      t.keys[h] = VX_EMPTY_KEY;
      t.counts[h] = 0;
    }
  }
}

__global__ void __launch_bounds__(256) vx_table_reset_kernel(const VxSlot* __restrict__ slots) {
  const VxSlot& s = slots[blockIdx.y >> 1];
  const VxTable& t = (blockIdx.y & 1u) ? s.target : s.input;
  const uint32_t used = *t.used_count;
  for (uint32_t u = blockIdx.x * blockDim.x + threadIdx.x; u < used; u += gridDim.x * blockDim.x) {
    const uint32_t h = t.used[u];
    t.keys[h] = VX_EMPTY_KEY;
    t.counts[h] = 0;
  }
}

// bit b of a 32-voxel ballot = voxel L0 + b.  pack() puts element 8n in bit 7 of byte n: reverse the
// bits of every byte = reverse the word, then swap its bytes back.
__device__ __forceinline__ uint32_t msb_first(uint32_t ballot) { return __byte_perm(__brev(ballot), 0u, 0x0123u); }

// K5a.  One warp per 32 consecutive voxels.
__global__ void __launch_bounds__(256) vx_emit_kernel(const VxSlot* __restrict__ slots, uint64_t nvox) {
  const VxSlot& s = slots[blockIdx.y];
  const uint32_t lane = threadIdx.x & 31u;
  const uint64_t warps = (uint64_t)gridDim.x * (blockDim.x >> 5);
  for (uint64_t L0 = ((uint64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * 32u; L0 < nvox; L0 += warps * 32u) {
    const uint64_t L = L0 + lane;
    const bool valid = L < nvox;
    unsigned long long bi = 0, bt = 0;
    if (valid) {
      bi = s.best_input[L];
      bt = s.best_target[L];
      if (bi) s.best_input[L] = 0;
      if (bt) s.best_target[L] = 0;
      s.out_label[L] = bt ? (uint16_t)(0xFFFFu - (uint32_t)(bt & 0xFFFFull)) : (uint16_t)0;
    }
    const uint32_t label_in = bi ? (0xFFFFu - (uint32_t)(bi & 0xFFFFull)) : 0u;
    const uint32_t occ_word = __ballot_sync(0xFFFFFFFFu, bt != 0);
    const uint32_t bin_word = __ballot_sync(0xFFFFFFFFu, label_in > 0u);
    if (lane == 0) {
      s.occ[L0 >> 5] = occ_word;
      s.out_bin[L0 >> 5] = msb_first(bin_word);
    }
  }
}

// K5b.  isOccluded = occlusions_ > -1; isInvalid = invalid_ > -1 && invalid_ != own index, i.e. still
// alive after every pass and not occupied (VoxelGrid.cpp:65-73).
__global__ void __launch_bounds__(256) vx_pack_kernel(const VxSlot* __restrict__ slots, uint32_t words) {
  const VxSlot& s = slots[blockIdx.y];
  for (uint32_t w = blockIdx.x * blockDim.x + threadIdx.x; w < words; w += gridDim.x * blockDim.x) {
    const uint32_t occ = s.occ[w], occl = s.occl[w], alive = s.alive[w];
    s.out_occluded[w] = msb_first(occl);
    s.out_invalid[w] = msb_first(alive & ~occ);
  }
}

}  // namespace

cudaError_t vx_launch_fill(const VxFillItem* items, uint32_t n_items, uint32_t max_points, const float* ap,
                           const VxSlot* slots, VxGridGeom geom, VxFilterDev filter, cudaStream_t stream) {
  if (n_items == 0 || max_points == 0) return cudaSuccess;
  const uint32_t tiles = (max_points + kFillTile - 1) / kFillTile;
  for (uint32_t first = 0; first < n_items; first += 65535u) {
    const uint32_t n = n_items - first < 65535u ? n_items - first : 65535u;
    vx_fill_kernel<<<dim3(tiles, n), kFillThreads, 0, stream>>>(items + first, ap, slots, geom, filter);
  }
  return cudaGetLastError();
}

cudaError_t vx_launch_vote(const VxSlot* slots, uint32_t n_slots, int keep_tables, cudaStream_t stream) {
  if (n_slots == 0) return cudaSuccess;
  vx_vote_kernel<<<dim3(32, 2 * n_slots), 256, 0, stream>>>(slots, keep_tables);
  return cudaGetLastError();
}

cudaError_t vx_launch_table_reset(const VxSlot* slots, uint32_t n_slots, cudaStream_t stream) {
  if (n_slots == 0) return cudaSuccess;
  vx_table_reset_kernel<<<dim3(32, 2 * n_slots), 256, 0, stream>>>(slots);
  return cudaGetLastError();
}

cudaError_t vx_launch_emit(const VxSlot* slots, uint32_t n_slots, uint64_t nvox, cudaStream_t stream) {
  if (n_slots == 0) return cudaSuccess;
  const uint64_t warps_needed = (nvox + 31) / 32;
  uint32_t blocks = (uint32_t)((warps_needed + 7) / 8);
  if (blocks > 1024u) blocks = 1024u;
  vx_emit_kernel<<<dim3(blocks, n_slots), 256, 0, stream>>>(slots, nvox);
  return cudaGetLastError();
}

cudaError_t vx_launch_pack(const VxSlot* slots, uint32_t n_slots, uint64_t nvox, cudaStream_t stream) {
  if (n_slots == 0) return cudaSuccess;
  const uint32_t words = (uint32_t)((nvox + 31) / 32);
  uint32_t blocks = (words + 255) / 256;
  if (blocks > 256u) blocks = 256u;
  vx_pack_kernel<<<dim3(blocks, n_slots), 256, 0, stream>>>(slots, words);
  return cudaGetLastError();
}
