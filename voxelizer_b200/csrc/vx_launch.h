// Host-callable launch wrappers of the kernels (defined in vx_fill.cu / vx_visibility.cu).
#ifndef VOXELIZER_B200_VX_LAUNCH_H_
#define VOXELIZER_B200_VX_LAUNCH_H_

#include "vx_device.cuh"

// K1+K2: transform + filter + voxel key + (voxel,label) histogram insert, organised by scan: grid =
// (tiles, scans of the launch); every tile is loaded once and inserted into each frame that uses the scan.
cudaError_t vx_launch_fill(const VxScanWork* works, uint32_t n_works, uint32_t max_points, const VxScanUse* uses, const float* ap,
                           const VxSlot* slots, VxGridGeom geom, VxFilterDev filter, cudaStream_t stream);

// K3: clears the dense outputs, then one RED.MAX per histogram slot into the dense vote array best[L].
cudaError_t vx_launch_vote(const VxSlot* slots, uint32_t n_slots, uint64_t nvox, cudaStream_t stream);
// Frees tables that were kept for debugging.
cudaError_t vx_launch_table_reset(const VxSlot* slots, uint32_t n_slots, cudaStream_t stream);

// K5a: winning slots -> .label (uint16), .bin (packed, input grid), occupancy bitmap; frees the histogram
// slots (unless keep_tables) and resets best[].
cudaError_t vx_launch_emit(const VxSlot* slots, uint32_t n_slots, int keep_tables, cudaStream_t stream);

// K4: occlusion pass + invalid passes, one CTA per frame; every CTA claims one of the n_scratch scratch
// buffers (locks[i]: 0 free / 1 taken) -- n_scratch must be >= vx_visibility_resident_ctas().
cudaError_t vx_visibility_prepare();  // opt-in to large dynamic shared memory (once per process)
int vx_visibility_resident_ctas(int device);  // CTAs of the kernel the device holds at once
cudaError_t vx_launch_visibility(const VxSlot* slots, const VxVisFrame* frames, uint32_t n_slots,
                                 const float* endpoints, VxGridGeom geom, const VxVisScratch* scratch, uint32_t n_scratch,
                                 uint32_t* locks, int mode, int leave_room, cudaStream_t stream);
uint64_t vx_visibility_visits(VxGridGeom geom);  // visits per pass (must stay < 2^25)

// VoxelGrid::occludedBy as a public call: one ray over a frame's occupancy; visited = capacity x (i, j, k), result = {cells
// looked at, occluder (reference-internal index) or -1}.
cudaError_t vx_launch_trace_ray(const uint32_t* occ, VxGridGeom geom, int32_t i, int32_t j, int32_t k, const float* endpoint,
                                int32_t* visited, uint32_t capacity, int32_t* result, cudaStream_t stream);

// K5b: occl / alive bitmaps -> MSB-first packed .occluded / .invalid.
cudaError_t vx_launch_pack(const VxSlot* slots, uint32_t n_slots, uint64_t nvox, cudaStream_t stream);

#endif
