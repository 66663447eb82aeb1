/* voxelizer_b200 -- C ABI of the B200 backend for jbehley/voxelizer's gen_data hot path.
 *
 * The reference (single-threaded C++, no FFI of its own) runs this path behind plain classes:
 *   gen_data frame loop                 src/gen_data.cpp:62-142
 *   fillVoxelGrid                       src/data/voxelize_utils.cpp:173-203
 *   VoxelGrid::insert / clear           src/data/VoxelGrid.cpp:30-63
 *   VoxelGrid::updateOcclusions         src/data/VoxelGrid.cpp:98-171
 *   VoxelGrid::updateInvalid            src/data/VoxelGrid.cpp:173-233
 *   VoxelGrid::occludedBy               src/data/VoxelGrid.cpp:235-316
 *   saveVoxelGrid / pack                src/data/voxelize_utils.cpp:205-296
 * This library replaces exactly that path.  It sits BELOW the reference's Eigen pose algebra: the
 * caller computes, per scan of a frame's window, ap = anchor_pose.inverse() * pose
 * (voxelize_utils.cpp:183) and the endpoint (gen_data.cpp:114) on the host and passes the numbers in,
 * so the GPU never has to reproduce Eigen's 4x4 inverse.  The C++ facade in voxelizer_b200/host keeps
 * the reference's own API (VoxelGrid, fillVoxelGrid, saveVoxelGrid, Config, KittiReader) on top of it.
 *
 * Conventions: every entry point returns 0 on success or a negative vx_status; CUDA failures are
 * reported, never papered over -- there is NO CPU fallback.  Matrices are 16 floats, column-major
 * (Eigen::Matrix4f storage order).  The caller owns all host buffers; the library owns all device
 * memory.  One context per GPU; calls on one context must be serialised by the caller (one thread at a
 * time; the library's own worker thread behind vx_submit_frames is synchronised inside).
 * All outputs are byte-identical to the files the reference writes.
 */
#ifndef VOXELIZER_B200_H_
#define VOXELIZER_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct vx_ctx vx_ctx;

typedef enum vx_status {
  VX_OK = 0,
  VX_ERR_INVALID = -1,        /* bad argument */
  VX_ERR_CUDA = -2,           /* a CUDA runtime call or kernel failed; see vx_last_error */
  VX_ERR_NO_DEVICE = -3,      /* no usable CUDA device: the library does not run on the CPU */
  VX_ERR_UNKNOWN_SCAN = -4,   /* a frame references a scan id that is not resident */
  VX_ERR_TABLE_OVERFLOW = -5, /* histogram table could not be grown any further */
  VX_ERR_UNSUPPORTED = -6     /* grid too large for the traversal's visit counter (> 2^25 visits) */
} vx_status;

/* Grid geometry = the arguments of VoxelGrid::initialize(resolution, min, max)
 * (VoxelGrid.cpp:5-28); sizes and offset are derived inside with the reference's arithmetic. */
typedef struct vx_grid_desc {
  float resolution;    /* Config::voxelSize  */
  float min_extent[3]; /* Config::minExtent  */
  float max_extent[3]; /* Config::maxExtent  */
} vx_grid_desc;

/* Point filters of fillVoxelGrid (voxelize_utils.cpp:175-201). */
typedef struct vx_filter_desc {
  float min_range;                 /* Config::minRange */
  float max_range;                 /* Config::maxRange */
  int32_t hidecar;                 /* Config::hidecar (ego-vehicle box -2 < x < 3, |y| < 2) */
  const uint32_t* filtered_labels; /* Config::filteredLabels ("ignore") */
  uint32_t n_filtered;
  const uint32_t* join_pairs; /* Config::joinedLabels flattened as (key, member) pairs in map order */
  uint32_t n_join_pairs;
} vx_filter_desc;

/* Tuning knobs; zero-initialise for defaults. */
typedef struct vx_options {
  uint32_t frames_in_flight;     /* frames per internal batch = one visibility launch, one CTA per frame
                                    (default: 4 x the CTAs the device holds at once, memory permitting) */
  uint32_t table_log2_capacity;  /* initial (voxel,label) histogram slots per frame, log2 (default 18; grown x4
                                    and the batch redone when a frame overflows it) */
  void* stream;                  /* cudaStream_t to run on (e.g. the caller's framework stream so that its
                                    events bracket the work); NULL = a private non-blocking stream */
  uint32_t debug;                /* 1: keep histogram tables after a batch for vx_debug_dump */
  uint32_t table_sets;           /* frames per fill -> vote -> emit launch; frames this far apart share one set of
                                    histogram tables (default 128) */
  uint32_t vis_mode;             /* shape of the traversal kernel: 0 = chosen per batch (few frames: one 512-thread CTA per SM and
                                    whole-face waves, short time per frame; many frames: 7 x 128-thread CTAs per SM, most frames per
                                    second), 1 = always the many-frames shape, 2 = always the few-frames shape, 3 = the shape in between (4 x 256-thread
                                    CTAs per SM).  Same bytes either way. */
  uint32_t occlusion_labels;     /* 1: frames may ask for VoxelGrid::insertOcclusionLabels (VX_FRAME_INSERT_OCCLUSION_LABELS); costs 8 bytes per
                                    voxel and frame of a batch, and 4 per visit of the traversal's scratch */
} vx_options;

/* One target frame = one iteration of gen_data.cpp:62-142 that passed the "labelled" gate.
 * prior_* describe priorPoints (the last one is the target scan), past_* describe pastPoints.
 * Prior scans are inserted into BOTH grids, past scans into the target grid only
 * (gen_data.cpp:96-98).  endpoints[i] = (anchor^-1 * pose_past[i]).col(3).head(3), one
 * VoxelGrid::updateInvalid pass each, in this order (gen_data.cpp:113-116).
 * Output pointers may be NULL to skip that product; sizes are vx_output_bytes(). */
typedef struct vx_frame_desc {
  uint32_t n_prior;
  const uint32_t* prior_ids;
  const float* ap_prior; /* n_prior x 16 */
  uint32_t n_past;
  const uint32_t* past_ids;
  const float* ap_past;   /* n_past x 16 */
  uint32_t n_endpoints;   /* updateInvalid calls; gen_data: one per past scan (= n_past) */
  const float* endpoints; /* n_endpoints x 3 */
  uint8_t* out_bin;       /* <name>.bin       nvox/8 bytes  (input grid, majority label > 0) */
  uint16_t* out_label;    /* <name>.label     nvox uint16   (target grid majority label)     */
  uint8_t* out_occluded;  /* <name>.occluded  nvox/8 bytes                                   */
  uint8_t* out_invalid;   /* <name>.invalid   nvox/8 bytes                                   */
  uint32_t flags;         /* vx_frame_flags */
  uint32_t reserved;
} vx_frame_desc;

/* VX_FRAME_INSERT_OCCLUSION_LABELS: run VoxelGrid::insertOcclusionLabels (VoxelGrid.cpp:75-96; the calls gen_data.cpp:108-109
 * and Mainframe.cpp:303-304 have commented out) on the TARGET grid between updateOcclusions and the updateInvalid passes:
 * every voxel that was empty takes count and labels of the first occupied voxel above it when only occluded empty voxels
 * lie between; the filled voxels are occupied for the passes that follow.  Needs vx_options.occlusion_labels.
 * VX_FRAME_INPUT_IS_TARGET: the input grid holds the same points as the target grid (a single VoxelGrid behind the
 * facade): `.bin` is then packed from the target grid's labels, filled ones included. */
typedef enum vx_frame_flags { VX_FRAME_INSERT_OCCLUSION_LABELS = 1, VX_FRAME_INPUT_IS_TARGET = 2 } vx_frame_flags;

typedef struct vx_grid_info {
  uint32_t size[3]; /* Sx, Sy, Sz */
  float offset[3];
  float resolution;
  uint64_t num_voxels;
  uint64_t packed_bytes; /* num_voxels / 8 */
  uint64_t label_bytes;  /* num_voxels * 2 */
} vx_grid_info;

/* Accumulated device time per stage (CUDA events on the library's stream) since the last reset. */
typedef struct vx_stage_times {
  double fill_ms;       /* K1+K2: transform, filter, key, histogram insert */
  double vote_ms;       /* K3: per-voxel majority vote                      */
  double emit_ms;       /* K5: label / .bin / occupancy emit + table reset  */
  double visibility_ms; /* K4: occlusion + invalid passes                   */
  double pack_ms;       /* K5: .occluded / .invalid bit-pack                */
  double total_ms;      /* whole batches, first kernel to last copy         */
  uint64_t frames;
  uint64_t points_read;   /* window points streamed by the fill kernel */
  uint64_t fill_launches; /* kernel launches per stage */
  uint64_t vote_launches;
  uint64_t emit_launches;
  uint64_t visibility_launches;
  uint64_t pack_launches;
  uint64_t rays_cast; /* occludedBy calls (all passes) */
  uint64_t dda_steps; /* cells traversed by those rays */
  /* K4 anatomy, summed over frames: waves run, waves that needed the in-wave resolution, resolution
   * sweeps, and SM cycles as seen by thread 0 of each frame's CTA: [0..3] per phase (snapshot, resolve,
   * trace, apply); [4..6] developer statistics of builds with
   * -DVX_TRACE_STATS (sum over waves of the longest ray in steps, step-loop iterations and attention events of thread 0),
   * otherwise zero; [7] reserved (zero). */
  uint64_t vis_waves;
  uint64_t vis_resolve_waves;
  uint64_t vis_rounds;
  uint64_t vis_cycles[8];
} vx_stage_times;

const char* vx_last_global_error(void); /* for failures before a context exists */
const char* vx_last_error(const vx_ctx* ctx);

int vx_device_count(void);
int vx_create(int device, const vx_grid_desc* grid, const vx_filter_desc* filter, const vx_options* options, vx_ctx** out);
void vx_destroy(vx_ctx* ctx);
int vx_grid_info_get(const vx_ctx* ctx, vx_grid_info* out);
uint32_t vx_frames_in_flight(const vx_ctx* ctx); /* frames one internal batch holds */
uint32_t vx_batches_in_flight(const vx_ctx* ctx); /* batches the library keeps in flight at a time (its lanes): a caller that
                                                     streams jobs keeps this many vx_submit_frames tickets outstanding */

/* Pinned host memory helpers for the output ring / scan staging. */
void* vx_host_alloc(size_t bytes);
void vx_host_free(void* p);

/* Device-resident scan cache (replaces KittiReader's pointsCache_/labelCache_, KittiReader.cpp:73-142).
 * xyzr: n x 4 float32 exactly as in a KITTI velodyne .bin (KittiReader.cpp:145-171); labels: n raw
 * uint32 words of the .label file -- the library applies the reader's "& 0xFFFF" (KittiReader.cpp:189).
 * The copy is enqueued on an upload stream of the context: from pageable memory it has completed when the
 * call returns, from pinned memory (vx_host_alloc) it may still be in flight -- vx_process_frames waits (on the
 * device) for the scans its frames read and for nothing else, so frames can be voxelized while later scans
 * are still arriving; call vx_sync before overwriting a pinned source buffer. */
int vx_upload_scan(vx_ctx* ctx, uint32_t scan_id, const float* xyzr, const uint32_t* labels, uint32_t n);
/* Completion of uploads, for callers that recycle a small pinned staging ring: vx_upload_ticket returns the number
 * of uploads enqueued so far, vx_upload_wait blocks until that many have arrived on the device. */
uint64_t vx_upload_ticket(const vx_ctx* ctx);
int vx_upload_wait(vx_ctx* ctx, uint64_t ticket);
int vx_evict_scan(vx_ctx* ctx, uint32_t scan_id);
int vx_evict_all(vx_ctx* ctx);

/* Runs n frames (any n; processed in batches of frames_in_flight) and returns after all outputs landed:
 * vx_submit_frames followed by vx_wait. */
int vx_process_frames(vx_ctx* ctx, const vx_frame_desc* frames, uint32_t n);

/* Asynchronous form (vx_submit_frame -> ticket -> vx_wait).  vx_submit_frames copies the descriptors and
 * the id / matrix / endpoint arrays they point to and returns at once; a worker thread of the context runs the jobs in
 * submission order.  Until vx_wait(ticket) has returned the caller must leave the job's OUTPUT buffers alone and must
 * not evict or re-upload a scan id the job reads; uploading OTHER scans meanwhile is the point of this call (the next
 * sequence crosses PCIe while this one is traversed).  vx_wait returns the job's status (its message through
 * vx_last_error); ticket 0 waits for everything submitted so far and returns the first failure.  vx_last_error,
 * vx_stage_times_get and vx_debug_dump describe finished jobs only: call them after vx_wait. */
int vx_submit_frames(vx_ctx* ctx, const vx_frame_desc* frames, uint32_t n, uint64_t* ticket);
int vx_wait(vx_ctx* ctx, uint64_t ticket);
int vx_sync(vx_ctx* ctx); /* every job, every upload, the stream */

int vx_stage_times_get(vx_ctx* ctx, vx_stage_times* out);
int vx_stage_times_reset(vx_ctx* ctx);

/* Parity hooks: internals of frame `frame_index` of the LAST vx_process_frames call (must have been
 * the last batch, i.e. n <= frames_in_flight).  Index order is the output order
 * L = (x*Sy + y)*Sz + z.  dst may be NULL to query the required element count (returned via *count). */
typedef enum vx_dump_kind {
  VX_DUMP_HIST_TARGET = 0, /* uint32 triples (L, label, count), unordered         */
  VX_DUMP_HIST_INPUT = 1,
  VX_DUMP_OCCUPANCY = 2,   /* uint8 per voxel, target grid count > 0               */
  VX_DUMP_OCCLUDED = 3,    /* uint8 per voxel, isOccluded                          */
  VX_DUMP_INVALID = 4,     /* uint8 per voxel, isInvalid                           */
  VX_DUMP_FILL_SOURCE = 5  /* uint32 per voxel: L of the voxel insertOcclusionLabels copied from, 0xFFFFFFFF if none */
} vx_dump_kind;
int vx_debug_dump(vx_ctx* ctx, uint32_t frame_index, int kind, void* dst, uint64_t* count);

/* VoxelGrid::occludedBy(i, j, k, endpoint, visited*) as the PUBLIC method (VoxelGrid.h:89-90, VoxelGrid.cpp:235-316) on the
 * occupancy of frame `frame_index` of the last finished job (same restriction as vx_debug_dump): one ray from the centre
 * of voxel (i,j,k) towards `endpoint`.  *occluder = the reference-internal index i + j*Sx + k*Sx*Sy of the occupied voxel
 * that stops it, or -1; visited_ijk receives up to visited_capacity (i,j,k) triples of the cells the reference loop looks
 * at, in order (the occupied one included), *n_visited their full number.  visited_ijk / n_visited may be NULL. */
int vx_trace_ray(vx_ctx* ctx, uint32_t frame_index, int32_t i, int32_t j, int32_t k, const float* endpoint, int32_t* visited_ijk,
                 uint32_t visited_capacity, uint32_t* n_visited, int32_t* occluder);

#ifdef __cplusplus
}
#endif
#endif /* VOXELIZER_B200_H_ */
