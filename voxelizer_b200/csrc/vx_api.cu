// C ABI of the B200 backend (include/voxelizer_b200.h): context, device-resident scan cache, frame
// slots and the batch pipeline  fill -> vote -> emit -> visibility -> pack -> copy-out.
// Host-side only; the kernels live in vx_fill.cu / vx_visibility.cu.  No CPU fallback anywhere: if
// CUDA is unusable every entry point reports VX_ERR_NO_DEVICE / VX_ERR_CUDA.
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <algorithm>
#include <condition_variable>
#include <cstring>
#include <deque>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

#include <nvtx3/nvToolsExt.h>

#include "../../include/voxelizer_b200.h"
#include "vx_launch.h"

namespace {

std::string g_global_error;

struct DeviceScan {
  float4* points = nullptr;
  uint32_t* labels = nullptr;
  uint32_t n = 0;
  uint64_t seq = 0;  // upload sequence number (uploads run on their own stream, see UploadMark)
};

// Uploads are asynchronous on a copy stream of their own; every few of them an event is recorded there.  A
// fill launch waits for the first mark at or after the newest scan it reads, so the first frames of a job are
// voxelized while the rest of the sequence is still crossing PCIe.
struct UploadMark {
  uint64_t seq;
  cudaEvent_t ev;
};

constexpr uint32_t kUploadsPerMark = 16;

// Histogram + winner tables: shared by the frames of a batch that are `table_sets` apart.
struct TableSet {
  VxTable input{}, target{};
};

// Per-frame state that outlives fill -> vote -> emit: bitmaps and the four output images.
struct FrameRec {
  VxSlot dev{};  // device pointers (tables filled in from the frame's table set)
};

struct ScratchMem {
  VxVisScratch dev{};
};

enum { EV_BEGIN = 0, EV_VIS_BEGIN, EV_VIS, EV_PACK, EV_END, EV_COUNT };

// per-slot counter block (uint32 words): used/overflow of both tables, then two 64-bit statistics
constexpr size_t kCounterWords = 4 + 2 * VX_SLOT_STATS + 2;  // 38 words per slot

// One asynchronous vx_submit_frames call: the descriptors and every array they point to are copied, so the
// caller's memory is free again when the call returns (the OUTPUT buffers are the caller's until vx_wait).
struct Job {
  uint64_t ticket = 0;
  std::vector<vx_frame_desc> frames;
  std::vector<uint32_t> ids;
  std::vector<float> ap, ep;
  int rc = 0;
  std::string error;
};

// One batch in flight.  The lanes take turns, so that the next batches (of this or the next jobs) are filled, voted and
// already tracing while the slowest frames of the previous one finish: a batch ends with its slowest frame (~1.2 x the mean
// frame time), and CTAs of the next launch take the slots the finished frames leave.  Two lanes by default; up to kMaxLanes
// (VX_B200_LANES, a developer switch: measured, the traversal slots are full with two).
constexpr uint32_t kMaxLanes = 4;
struct Lane {
  cudaStream_t stream = nullptr;
  cudaStream_t fill_stream = nullptr;  // high-priority stream of the fill -> vote -> emit launches (null: they run on `stream`)
  cudaEvent_t ev_staged = nullptr, ev_filled = nullptr;  // hand-over between the two streams
  cudaEvent_t ev[5]{};               // EV_BEGIN .. EV_END
  std::vector<cudaEvent_t> ev_sub;   // 4 per fill -> vote -> emit launch
  unsigned char* h_stage = nullptr;  // batch staging (items, ap, endpoints, vis frames): pinned host + device mirror
  unsigned char* d_stage = nullptr;
  size_t stage_cap = 0;
  uint32_t slot_base = 0;            // its frames use slots slot_base .. slot_base + nb - 1
  // the batch
  bool busy = false;
  uint64_t seq = 0;                  // enqueue order
  const vx_frame_desc* frames = nullptr;
  uint32_t nb = 0, n_sub = 0;
  uint64_t points = 0, fill_launches = 0;
  Job* job = nullptr;                // the job the batch belongs to
  std::unique_ptr<Job> owned;        // ... owned here if this is the job's last batch
};

// developer timeline (VX_B200_TRACE=<file>): one line per host-side event, CLOCK_REALTIME ns (the device's %globaltimer of
// VX_B200_FRAME_STATS runs on the same clock)
void trace_line(const char* what, unsigned long long a, unsigned long long b, const std::string& rest = std::string()) {
  const char* path = std::getenv("VX_B200_TRACE");  // (read every time: a tool may switch files between contexts)
  if (!path) return;
  const unsigned long long now = (unsigned long long)std::chrono::duration_cast<std::chrono::nanoseconds>(
                                     std::chrono::system_clock::now().time_since_epoch()).count();
  if (FILE* fp = std::fopen(path, "a")) {
    std::fprintf(fp, "%s %llu %llu %llu %s\n", what, now, a, b, rest.c_str());
    std::fclose(fp);
  }
}

thread_local std::string* tl_error_sink = nullptr;  // set on the worker thread: errors belong to the job, not the context

struct NvtxRange {  // per-stage ranges on the host side of a batch (Stopwatch tic/toc -> NVTX)
  explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
  ~NvtxRange() { nvtxRangePop(); }
};

}  // namespace

struct vx_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  VxGridGeom geom{};
  uint64_t nvox = 0;
  uint64_t visits = 0;  // cell visits per visibility pass
  uint32_t words = 0;
  float min_extent[3]{}, max_extent[3]{};

  VxFilterDev filter{};
  uint32_t* d_reject_bits = nullptr;

  std::unordered_map<uint32_t, DeviceScan> scans;
  cudaMemPool_t scan_pool = nullptr;  // private stream-ordered pool of the scan cache

  uint32_t max_slots = 0;   // frame records: depth x lane_cap (a lane holds one batch = one visibility launch)
  uint32_t table_sets = 0;  // frames per fill -> vote -> emit launch
  uint32_t log2_target = 18, log2_input = 15;
  bool occlusion_labels = false;  // vx_options.occlusion_labels: slots and scratch carry the arrays insertOcclusionLabels needs
  int vis_mode = 0;  // 0 by batch size, 1 throughput shape, 2 latency shape (vx_options.vis_mode)
  bool keep_tables = false;
  bool tables_dirty = false;  // tables hold the previous batch (debug mode)
  uint32_t dirty_slots = 0;
  std::vector<TableSet> tables;
  std::vector<FrameRec> slots;
  std::vector<ScratchMem> scratch;  // one per resident visibility CTA
  std::vector<void*> slot_slabs, scratch_slabs;  // one cudaMalloc per growth step, carved into the records above
  VxVisScratch* d_scratch = nullptr;
  uint32_t* d_locks = nullptr;
  VxSlot* d_slots = nullptr;
  uint32_t* d_counters = nullptr;  // max_slots * kCounterWords
  uint32_t* h_counters = nullptr;  // pinned mirror

  // copy streams: uploads, and the outputs that are final before the traversal starts
  cudaStream_t copy_in = nullptr, copy_out = nullptr;
  cudaEvent_t ev_copy_out = nullptr;
  uint64_t upload_seq = 0, marked_seq = 0, done_seq = 0;
  std::vector<UploadMark> marks;      // in upload order; completed ones are recycled
  std::vector<cudaEvent_t> ev_pool;   // timing-less events

  Lane lanes[kMaxLanes];    // lanes[0].stream == stream
  bool leave_room = false;  // see vx_launch_visibility
  uint32_t depth = 1;       // batches in flight at a time (2 unless the tables are kept for debugging or memory is short)
  uint32_t lane_cap = 0;    // frames per batch
  uint32_t next_lane = 0;
  uint64_t batch_seq = 0;
  std::vector<uint8_t> slot_ready;  // frame record allocated?

  vx_stage_times times{};
  uint32_t last_batch = 0, last_base = 0;
  std::string error;

  // Asynchronous jobs (vx_submit_frames / vx_wait): one worker thread runs them in submission order.  `mu` guards
  // what the worker shares with the caller's upload / evict calls: the scan map, the upload marks and counters.
  std::mutex mu;
  std::mutex qmu;
  std::condition_variable qcv, dcv;
  std::deque<std::unique_ptr<Job>> queue;
  std::map<uint64_t, std::pair<int, std::string>> failed;  // tickets that did not end with VX_OK
  uint64_t next_ticket = 1, finished_ticket = 0;
  bool stop = false;
  std::thread worker;
};

namespace {

int fail(vx_ctx* c, int code, const std::string& msg) {
  if (tl_error_sink) *tl_error_sink = msg;
  else if (c) c->error = msg;
  else g_global_error = msg;
  return code;
}

#define VX_CUDA(ctx, call)                                                                         \
  do {                                                                                             \
    cudaError_t e__ = (call);                                                                      \
    if (e__ != cudaSuccess)                                                                        \
      return fail((ctx), VX_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__));       \
  } while (0)

size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// VoxelGrid::initialize (VoxelGrid.cpp:8-18): sizes by float ceil, offset with the double-precision
// "0.5 *" of the reference.  Host code (x86-64, no FMA contraction of separate statements).
void derive_geometry(const vx_grid_desc& d, VxGridGeom& g) {
  g.res = d.resolution;
  volatile float sx = std::ceil((d.max_extent[0] - d.min_extent[0]) / d.resolution);
  volatile float sy = std::ceil((d.max_extent[1] - d.min_extent[1]) / d.resolution);
  volatile float sz = std::ceil((d.max_extent[2] - d.min_extent[2]) / d.resolution);
  const uint32_t n[3] = {(uint32_t)sx, (uint32_t)sy, (uint32_t)sz};
  g.nx = (int32_t)n[0];
  g.ny = (int32_t)n[1];
  g.nz = (int32_t)n[2];
  for (int a = 0; a < 3; ++a) {
    volatile float span = d.max_extent[a] - d.min_extent[a];
    volatile float cover = (float)n[a] * d.resolution;
    volatile float slack = cover - span;
    g.off[a] = (float)((double)d.min_extent[a] - 0.5 * (double)slack);
  }
}

size_t table_bytes(uint32_t log2_cap) {
  const size_t cap = (size_t)1 << log2_cap;
  return align_up(cap * 8, 256) + align_up(cap * 4, 256);
}

size_t table_set_bytes(const vx_ctx* c) { return table_bytes(c->log2_target) + table_bytes(c->log2_input) + 2 * align_up(c->nvox * 8, 256); }

size_t frame_bytes(const vx_ctx* c) {
  size_t b = 6 * align_up((size_t)c->words * 4, 256) + align_up(c->nvox * 2, 256);
  if (c->occlusion_labels) b += align_up((size_t)c->words * 4, 256) + 2 * align_up(c->nvox * 4, 256);
  return b;
}

size_t scratch_bytes(const vx_ctx* c) {
  return align_up(c->nvox * 4, 256) + align_up((c->visits / 32 + 1) * 4, 256) + (c->occlusion_labels ? align_up((c->visits + 32) * 4, 256) : 0);
}

struct Carver {
  unsigned char* p;
  unsigned char* take(size_t bytes) {
    unsigned char* r = p;
    p += align_up(bytes, 256);
    return r;
  }
};

void carve_table_set(vx_ctx* c, TableSet& t, unsigned char* arena) {
  Carver cv{arena};
  auto carve = [&](VxTable& tb, uint32_t log2_cap) {
    const size_t cap = (size_t)1 << log2_cap;
    tb.keys = reinterpret_cast<unsigned long long*>(cv.take(cap * 8));
    tb.counts = reinterpret_cast<uint32_t*>(cv.take(cap * 4));
    tb.best = reinterpret_cast<unsigned long long*>(cv.take(c->nvox * 8));
    tb.log2_cap = log2_cap;
  };
  carve(t.target, c->log2_target);
  carve(t.input, c->log2_input);
}

void carve_frame(vx_ctx* c, uint32_t index, unsigned char* arena) {
  FrameRec& s = c->slots[index];
  Carver cv{arena};
  // (the histogram tables a frame uses depend on its place IN ITS BATCH, whichever lane runs the batch)
  const TableSet& t = c->tables[(index % c->lane_cap) % c->table_sets];
  uint32_t* cnt = c->d_counters + (size_t)index * kCounterWords;
  s.dev.target = t.target;
  s.dev.target.overflow = cnt + 1;
  s.dev.input = t.input;
  s.dev.input.overflow = cnt + 3;
  s.dev.counters = reinterpret_cast<unsigned long long*>(cnt + 4);
  s.dev.occ = reinterpret_cast<uint32_t*>(cv.take((size_t)c->words * 4));
  s.dev.occl = reinterpret_cast<uint32_t*>(cv.take((size_t)c->words * 4));
  s.dev.alive = reinterpret_cast<uint32_t*>(cv.take((size_t)c->words * 4));
  s.dev.out_label = reinterpret_cast<uint16_t*>(cv.take(c->nvox * 2));
  s.dev.out_bin = reinterpret_cast<uint32_t*>(cv.take((size_t)c->words * 4));
  s.dev.out_occluded = reinterpret_cast<uint32_t*>(cv.take((size_t)c->words * 4));
  s.dev.out_invalid = reinterpret_cast<uint32_t*>(cv.take((size_t)c->words * 4));
  s.dev.occ_filled = s.dev.occluder = s.dev.fill_source = nullptr;
  if (c->occlusion_labels) {
    s.dev.occ_filled = reinterpret_cast<uint32_t*>(cv.take((size_t)c->words * 4));
    s.dev.occluder = reinterpret_cast<uint32_t*>(cv.take(c->nvox * 4));
    s.dev.fill_source = reinterpret_cast<uint32_t*>(cv.take(c->nvox * 4));
  }
}

void free_slots(vx_ctx* c) {
  for (void* p : c->slot_slabs) cudaFree(p);
  c->slot_slabs.clear();
  c->slots.clear();
  c->slot_ready.clear();
  c->tables.clear();
  c->tables_dirty = false;
}

// One cudaMalloc per growth step (a whole sequence needs ~900 frame records: one call, not 900).  Frame records of the slots
// base .. base + n - 1 (a lane's range), and one table set per frame of a fill group.
int ensure_slots(vx_ctx* c, uint32_t base, uint32_t n) {
  const uint32_t want_sets = n < c->table_sets ? n : c->table_sets;
  if (c->tables.size() < want_sets) {
    const size_t old = c->tables.size(), add = want_sets - old, each = table_set_bytes(c);
    void* slab = nullptr;
    VX_CUDA(c, cudaMalloc(&slab, add * each));
    c->slot_slabs.push_back(slab);
    VX_CUDA(c, cudaMemsetAsync(slab, 0, add * each, c->stream));  // all-zero = empty (vote / emit leave them empty again)
    VX_CUDA(c, cudaStreamSynchronize(c->stream));
    c->tables.resize(want_sets);
    for (size_t i = 0; i < add; ++i) carve_table_set(c, c->tables[old + i], static_cast<unsigned char*>(slab) + i * each);
    // (records carved earlier keep pointing at their set: index % table_sets is stable while table_sets is)
  }
  if (c->slots.size() < c->max_slots) {
    c->slots.resize(c->max_slots);
    c->slot_ready.assign(c->max_slots, 0);
  }
  uint32_t first = base;
  while (first < base + n && c->slot_ready[first]) ++first;
  if (first == base + n) return VX_OK;
  const size_t add = base + n - first, each = frame_bytes(c);
  void* slab = nullptr;
  VX_CUDA(c, cudaMalloc(&slab, add * each));
  c->slot_slabs.push_back(slab);
  std::vector<VxSlot> host(add);
  for (size_t i = 0; i < add; ++i) {
    carve_frame(c, (uint32_t)(first + i), static_cast<unsigned char*>(slab) + i * each);
    c->slot_ready[first + i] = 1;
    host[i] = c->slots[first + i].dev;
  }
  VX_CUDA(c, cudaMemcpy(c->d_slots + first, host.data(), add * sizeof(VxSlot), cudaMemcpyHostToDevice));
  return VX_OK;
}

// scratch buffers of the visibility kernel: one per CTA the device can hold (at most n_frames)
int ensure_scratch(vx_ctx* c, uint32_t n_frames) {
  uint32_t want = (uint32_t)vx_visibility_resident_ctas(c->device);
  if (want == 0) return fail(c, VX_ERR_CUDA, "the visibility kernel does not fit on this device");
  if (!c->d_scratch) {
    VX_CUDA(c, cudaMalloc(reinterpret_cast<void**>(&c->d_scratch), (size_t)want * sizeof(VxVisScratch)));
    VX_CUDA(c, cudaMalloc(reinterpret_cast<void**>(&c->d_locks), (size_t)want * sizeof(uint32_t)));
    VX_CUDA(c, cudaMemset(c->d_locks, 0, (size_t)want * sizeof(uint32_t)));
  }
  // one batch at a time: at most n_frames CTAs exist; two lanes: CTAs of both launches share the SMs, up to `want` are resident
  if (c->depth == 1 && want > n_frames) want = n_frames;
  if (c->scratch.size() >= want) return VX_OK;
  const size_t old = c->scratch.size(), each = scratch_bytes(c);
  void* slab = nullptr;
  VX_CUDA(c, cudaMalloc(&slab, (want - old) * each));
  c->scratch_slabs.push_back(slab);
  c->scratch.resize(want);
  for (size_t i = old; i < want; ++i) {
    Carver cv{static_cast<unsigned char*>(slab) + (i - old) * each};
    c->scratch[i].dev.memo = reinterpret_cast<uint32_t*>(cv.take(c->nvox * 4));
    c->scratch[i].dev.hitbits = reinterpret_cast<uint32_t*>(cv.take((c->visits / 32 + 1) * 4));
    c->scratch[i].dev.hitcell = c->occlusion_labels ? reinterpret_cast<uint32_t*>(cv.take((c->visits + 32) * 4)) : nullptr;
  }
  std::vector<VxVisScratch> host(want);
  for (size_t i = 0; i < want; ++i) host[i] = c->scratch[i].dev;
  VX_CUDA(c, cudaMemcpy(c->d_scratch, host.data(), (size_t)want * sizeof(VxVisScratch), cudaMemcpyHostToDevice));
  return VX_OK;
}

int ensure_stage(vx_ctx* c, Lane& ln, size_t bytes) {
  if (bytes <= ln.stage_cap) return VX_OK;
  if (ln.h_stage) cudaFreeHost(ln.h_stage);
  if (ln.d_stage) cudaFree(ln.d_stage);
  ln.h_stage = nullptr;
  ln.d_stage = nullptr;
  ln.stage_cap = 0;
  const size_t cap = align_up(bytes * 2, 4096);
  VX_CUDA(c, cudaMallocHost(reinterpret_cast<void**>(&ln.h_stage), cap));
  VX_CUDA(c, cudaMalloc(reinterpret_cast<void**>(&ln.d_stage), cap));
  ln.stage_cap = cap;
  return VX_OK;
}

int mark_uploads(vx_ctx* c) {
  if (c->marked_seq == c->upload_seq) return VX_OK;
  cudaEvent_t e = nullptr;
  if (!c->ev_pool.empty()) {
    e = c->ev_pool.back();
    c->ev_pool.pop_back();
  } else {
    VX_CUDA(c, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  }
  VX_CUDA(c, cudaEventRecord(e, c->copy_in));
  c->marks.push_back(UploadMark{c->upload_seq, e});
  c->marked_seq = c->upload_seq;
  return VX_OK;
}

void prune_marks(vx_ctx* c) {
  size_t k = 0;
  while (k < c->marks.size() && cudaEventQuery(c->marks[k].ev) == cudaSuccess) {
    c->done_seq = c->marks[k].seq;
    c->ev_pool.push_back(c->marks[k].ev);
    ++k;
  }
  if (k) c->marks.erase(c->marks.begin(), c->marks.begin() + (long)k);
  cudaGetLastError();  // cudaErrorNotReady of the query is not an error
}

// makes `st` wait until upload number `seq` has arrived; false if it already had
int wait_for_upload(vx_ctx* c, cudaStream_t st, uint64_t seq, bool* waited) {
  *waited = false;
  if (seq <= c->done_seq) return VX_OK;
  for (const UploadMark& m : c->marks)
    if (m.seq >= seq) {
      VX_CUDA(c, cudaStreamWaitEvent(st, m.ev, 0));
      *waited = true;
      return VX_OK;
    }
  return fail(c, VX_ERR_CUDA, "internal: upload without a mark");
}

// Enqueues one batch of <= lane_cap frames on a lane: stage, fill -> vote -> emit per group, ONE traversal launch, pack, the
// copies out.  Returns without waiting for the device (unless an output buffer is pageable memory: such a copy blocks).
// `prev` = the batch enqueued last if it is still in flight (null: none, or nothing of it may be waited for).
int enqueue_batch(vx_ctx* c, Lane& ln, Lane* prev, const vx_frame_desc* frames, uint32_t nb) {
  const uint32_t base = ln.slot_base;
  trace_line("enqueue_begin", c->batch_seq + 1, nb);
  int rc = ensure_slots(c, base, nb);
  if (rc != VX_OK) return rc;
  rc = ensure_scratch(c, nb);
  if (rc != VX_OK) return rc;

  // ---- stage the batch description (the scan map and the upload marks are shared with the caller's thread)
  NvtxRange nvtx_batch("vx:batch");
  nvtxRangePushA("vx:stage+enqueue");
  std::unique_lock<std::mutex> lk(c->mu);
  struct PopOnExit {
    bool armed = true;
    ~PopOnExit() { if (armed) nvtxRangePop(); }
  } nvtx_enqueue;
  size_t n_items = 0, n_ap = 0, n_ep = 0;
  for (uint32_t f = 0; f < nb; ++f) {
    n_items += frames[f].n_prior + frames[f].n_past;
    n_ap += frames[f].n_prior + frames[f].n_past;
    n_ep += frames[f].n_endpoints;
  }
  const size_t off_works = 0;
  const size_t off_uses = align_up(off_works + n_items * sizeof(VxScanWork), 256);
  const size_t off_ap = align_up(off_uses + n_items * sizeof(VxScanUse), 256);
  const size_t off_ep = align_up(off_ap + n_ap * 16 * sizeof(float), 256);
  const size_t off_vis = align_up(off_ep + n_ep * 3 * sizeof(float), 256);
  const size_t total = align_up(off_vis + nb * sizeof(VxVisFrame), 256);
  rc = ensure_stage(c, ln, total);
  if (rc != VX_OK) return rc;
  VxScanWork* works = reinterpret_cast<VxScanWork*>(ln.h_stage + off_works);
  VxScanUse* uses = reinterpret_cast<VxScanUse*>(ln.h_stage + off_uses);
  float* ap = reinterpret_cast<float*>(ln.h_stage + off_ap);
  float* ep = reinterpret_cast<float*>(ln.h_stage + off_ep);
  VxVisFrame* vis = reinterpret_cast<VxVisFrame*>(ln.h_stage + off_vis);

  // Fill work per group of `table_sets` frames, organised by scan: every scan a group touches becomes one
  // VxScanWork with the list of (frame, matrix, both) it goes into.
  struct Use {
    uint32_t scan_id, slot, ap_index, both;
  };
  const uint32_t n_sub = (nb + c->table_sets - 1) / c->table_sets;
  std::vector<size_t> work_begin(n_sub + 1, 0);
  std::vector<uint32_t> sub_max_points(n_sub, 0);
  std::vector<uint64_t> sub_need_seq(n_sub, 0);  // newest upload the group reads
  size_t wi = 0, ui = 0, ai = 0, ei = 0;
  uint64_t points = 0;
  std::vector<Use> group;
  for (uint32_t gi = 0; gi < n_sub; ++gi) {
    const uint32_t f0 = gi * c->table_sets, f1 = f0 + c->table_sets < nb ? f0 + c->table_sets : nb;
    group.clear();
    for (uint32_t f = f0; f < f1; ++f) {
      const vx_frame_desc& fr = frames[f];
      if ((fr.n_prior && (!fr.prior_ids || !fr.ap_prior)) || (fr.n_past && (!fr.past_ids || !fr.ap_past)) || (fr.n_endpoints && !fr.endpoints))
        return fail(c, VX_ERR_INVALID, "frame descriptor has null arrays");
      for (int part = 0; part < 2; ++part) {
        const uint32_t n = part == 0 ? fr.n_prior : fr.n_past;
        const uint32_t* ids = part == 0 ? fr.prior_ids : fr.past_ids;
        const float* aps = part == 0 ? fr.ap_prior : fr.ap_past;
        for (uint32_t s = 0; s < n; ++s) {
          std::memcpy(ap + 16 * ai, aps + 16 * (size_t)s, 16 * sizeof(float));
          group.push_back(Use{ids[s], f, (uint32_t)ai, part == 0 ? 1u : 0u});
          ++ai;
        }
      }
      if ((fr.flags & VX_FRAME_INSERT_OCCLUSION_LABELS) && !c->occlusion_labels)
        return fail(c, VX_ERR_INVALID, "VX_FRAME_INSERT_OCCLUSION_LABELS needs a context created with vx_options.occlusion_labels");
      vis[f].n_endpoints = fr.n_endpoints;
      vis[f].endpoint_offset = (uint32_t)ei;
      vis[f].flags = fr.flags;
      vis[f].pad = 0;
      if (fr.n_endpoints) std::memcpy(ep + 3 * ei, fr.endpoints, (size_t)fr.n_endpoints * 3 * sizeof(float));
      ei += fr.n_endpoints;
    }
    std::stable_sort(group.begin(), group.end(), [](const Use& a, const Use& b) { return a.scan_id < b.scan_id; });
    work_begin[gi] = wi;
    for (size_t k = 0; k < group.size();) {
      size_t k1 = k;
      while (k1 < group.size() && group[k1].scan_id == group[k].scan_id) ++k1;
      const auto it = c->scans.find(group[k].scan_id);
      if (it == c->scans.end()) return fail(c, VX_ERR_UNKNOWN_SCAN, "scan " + std::to_string(group[k].scan_id) + " is not resident");
      VxScanWork& w = works[wi++];
      w.points = it->second.points;
      w.labels = it->second.labels;
      w.n = it->second.n;
      w.use_begin = (uint32_t)ui;
      w.use_count = (uint32_t)(k1 - k);
      w.pad = 0;
      for (size_t q = k; q < k1; ++q) uses[ui++] = VxScanUse{group[q].slot, group[q].ap_index, group[q].both};
      if (w.n > sub_max_points[gi]) sub_max_points[gi] = w.n;
      if (it->second.seq > sub_need_seq[gi]) sub_need_seq[gi] = it->second.seq;
      points += (uint64_t)w.n * (k1 - k);
      k = k1;
    }
  }
  work_begin[n_sub] = wi;

  cudaStream_t st = ln.stream;
  if (c->tables_dirty) {  // debug mode kept the previous batch's tables
    VX_CUDA(c, vx_launch_table_reset(c->d_slots + c->last_base, c->dirty_slots, st));
    c->tables_dirty = false;
  }
  VX_CUDA(c, cudaMemcpyAsync(ln.d_stage, ln.h_stage, total, cudaMemcpyHostToDevice, st));
  VX_CUDA(c, cudaMemsetAsync(c->d_counters + (size_t)base * kCounterWords, 0, (size_t)nb * kCounterWords * sizeof(uint32_t), st));
  // frames `table_sets` apart share histogram tables, across lanes too: this batch's first fill waits for the last emit of the
  // batch enqueued before it (which waited for its own predecessor: the chain orders every batch in flight)
  if (prev && prev->busy && prev->n_sub) VX_CUDA(c, cudaStreamWaitEvent(st, prev->ev_sub[4 * (prev->n_sub - 1) + 3], 0));

  const VxScanWork* d_works = reinterpret_cast<const VxScanWork*>(ln.d_stage + off_works);
  const VxScanUse* d_uses = reinterpret_cast<const VxScanUse*>(ln.d_stage + off_uses);
  const float* d_ap = reinterpret_cast<const float*>(ln.d_stage + off_ap);
  const float* d_ep = reinterpret_cast<const float*>(ln.d_stage + off_ep);
  const VxVisFrame* d_vis = reinterpret_cast<const VxVisFrame*>(ln.d_stage + off_vis);

  // Uploads may still be in flight on the upload stream: each group waits for the newest scan it reads.
  int rc2 = mark_uploads(c);
  if (rc2 != VX_OK) return rc2;
  prune_marks(c);

  // fill -> vote -> emit in groups of `table_sets` frames (frames of different groups share tables)
  while (ln.ev_sub.size() < (size_t)n_sub * 4) {
    cudaEvent_t e;
    VX_CUDA(c, cudaEventCreate(&e));
    ln.ev_sub.push_back(e);
  }
  uint64_t fill_launches = 0;
  const uint32_t n_scratch = (uint32_t)c->scratch.size();
  VX_CUDA(c, cudaEventRecord(ln.ev[EV_BEGIN], st));
  // The short fill / vote / emit kernels run on a HIGH-PRIORITY stream of the lane: when a traversal CTA of an older batch
  // retires, the SM resources it leaves go to them before they go to the ~900 queued traversal CTAs of the batch in between.
  cudaStream_t fs = ln.fill_stream ? ln.fill_stream : st;
  if (fs != st) {
    VX_CUDA(c, cudaEventRecord(ln.ev_staged, st));
    VX_CUDA(c, cudaStreamWaitEvent(fs, ln.ev_staged, 0));
  }
  for (uint32_t g = 0; g < n_sub; ++g) {
    const uint32_t f0 = g * c->table_sets, f1 = f0 + c->table_sets < nb ? f0 + c->table_sets : nb;
    const uint32_t n_works = (uint32_t)(work_begin[g + 1] - work_begin[g]);
    bool waited = false;
    rc2 = wait_for_upload(c, fs, sub_need_seq[g], &waited);
    if (rc2 != VX_OK) return rc2;
    VX_CUDA(c, cudaEventRecord(ln.ev_sub[4 * g + 0], fs));
    VX_CUDA(c, vx_launch_fill(d_works + work_begin[g], n_works, sub_max_points[g], d_uses, d_ap, c->d_slots + base, c->geom, c->filter, fs));
    VX_CUDA(c, cudaEventRecord(ln.ev_sub[4 * g + 1], fs));
    VX_CUDA(c, vx_launch_vote(c->d_slots + base + f0, f1 - f0, c->nvox, fs));
    VX_CUDA(c, cudaEventRecord(ln.ev_sub[4 * g + 2], fs));
    VX_CUDA(c, vx_launch_emit(c->d_slots + base + f0, f1 - f0, c->keep_tables ? 1 : 0, fs));
    VX_CUDA(c, cudaEventRecord(ln.ev_sub[4 * g + 3], fs));
    fill_launches += (n_works + 65534) / 65535;
  }
  if (fs != st) {
    VX_CUDA(c, cudaEventRecord(ln.ev_filled, fs));
    VX_CUDA(c, cudaStreamWaitEvent(st, ln.ev_filled, 0));
  }
  // ONE visibility launch per batch.  (Measured: a launch per group on streams of their own, started as soon as the
  // group is filled, is slower -- every launch places its CTAs on the SMs in the same order, so 8 launches of 114
  // CTAs load 114 SMs with 7 CTAs and leave 34 idle; with groups of 148 it still lost 6 % to the single launch.)
  int leave_room = 0;  // VX_B200_LEAVE_ROOM=1 (developer switch, A/B): 6 traversal CTAs per SM while another batch is in flight / queued
  if (c->leave_room && c->depth >= 2) {
    bool queued = false;
    {
      std::lock_guard<std::mutex> qk(c->qmu);
      queued = !c->queue.empty();
    }
    leave_room = ((prev && prev->busy) || queued) ? 1 : 0;
  }
  VX_CUDA(c, cudaEventRecord(ln.ev[EV_VIS_BEGIN], st));
  VX_CUDA(c, vx_launch_visibility(c->d_slots + base, d_vis, nb, d_ep, c->geom, c->d_scratch, n_scratch, c->d_locks, c->vis_mode, leave_room, st));
  VX_CUDA(c, cudaEventRecord(ln.ev[EV_VIS], st));
  VX_CUDA(c, vx_launch_pack(c->d_slots + base, nb, c->nvox, st));
  VX_CUDA(c, cudaEventRecord(ln.ev[EV_PACK], st));
  lk.unlock();  // everything that reads the scan cache is enqueued; uploads / evictions of OTHER scans may go on
  nvtx_enqueue.armed = false;
  nvtxRangePop();
  NvtxRange nvtx_copy("vx:copy_out+wait");

  // Copy-out.  Everything above is enqueued by now: a copy into pageable memory blocks the host, never the device.
  // `.label` / `.bin` are final after a group's emit: they leave on the copy stream while the rays are traced.
  // ONE 2-D copy per run of frames whose host buffers are evenly spaced (an output ring) and whose frame records are too (carved
  // from one slab): a 909-frame batch used to queue 3636 copies, most of them behind the traversal kernel, and a stream takes
  // only ~1000 pending operations before the enqueueing call blocks -- measured: vx_submit_frames' worker sat in this loop until
  // the batch was nearly over, so batches with host outputs never overlapped (profiles/r02_e2e_timeline_before.txt).
  const size_t packed = (size_t)(c->nvox / 8);
  enum { OUT_LABEL = 0, OUT_BIN, OUT_OCCLUDED, OUT_INVALID };
  auto host_ptr = [&](uint32_t f, int kind) -> unsigned char* {
    const vx_frame_desc& fr = frames[f];
    void* p = kind == OUT_LABEL ? (void*)fr.out_label : kind == OUT_BIN ? (void*)fr.out_bin : kind == OUT_OCCLUDED ? (void*)fr.out_occluded : (void*)fr.out_invalid;
    return static_cast<unsigned char*>(p);
  };
  auto dev_ptr = [&](uint32_t f, int kind) -> const unsigned char* {
    const VxSlot& sl = c->slots[base + f].dev;
    const void* p = kind == OUT_LABEL ? (const void*)sl.out_label : kind == OUT_BIN ? (const void*)sl.out_bin : kind == OUT_OCCLUDED ? (const void*)sl.out_occluded : (const void*)sl.out_invalid;
    return static_cast<const unsigned char*>(p);
  };
  // frames [f0, f1) for which want(f) holds and the host buffer exists; returns the number of frames copied
  auto copy_out_kind = [&](uint32_t f0, uint32_t f1, int kind, cudaStream_t to, auto want, uint32_t* copied) -> cudaError_t {
    const size_t bytes = kind == OUT_LABEL ? (size_t)c->nvox * 2 : packed;
    const ptrdiff_t max_pitch = 0x7FFFFFFF;
    unsigned char *h0 = nullptr, *hprev = nullptr;
    const unsigned char *d0 = nullptr, *dprev = nullptr;
    ptrdiff_t hs = 0, ds = 0;
    uint32_t run = 0;
    auto flush = [&]() -> cudaError_t {
      cudaError_t e = cudaSuccess;
      if (run == 1) e = cudaMemcpyAsync(h0, d0, bytes, cudaMemcpyDeviceToHost, to);
      else if (run > 1) e = cudaMemcpy2DAsync(h0, (size_t)hs, d0, (size_t)ds, bytes, run, cudaMemcpyDeviceToHost, to);
      run = 0;
      return e;
    };
    for (uint32_t f = f0; f < f1; ++f) {
      unsigned char* hp = want(f) ? host_ptr(f, kind) : nullptr;
      if (!hp) {  // a frame that is left out ends the run (rows of a 2-D copy are consecutive)
        const cudaError_t e = flush();
        if (e != cudaSuccess) return e;
        continue;
      }
      const unsigned char* dp = dev_ptr(f, kind);
      ++*copied;
      if (run == 1) {
        hs = hp - hprev;
        ds = dp - dprev;
        if (hs >= (ptrdiff_t)bytes && ds >= (ptrdiff_t)bytes && hs <= max_pitch && ds <= max_pitch) {
          run = 2;
          hprev = hp;
          dprev = dp;
          continue;
        }
      } else if (run > 1 && hp - hprev == hs && dp - dprev == ds) {
        ++run;
        hprev = hp;
        dprev = dp;
        continue;
      }
      const cudaError_t e = flush();
      if (e != cudaSuccess) return e;
      h0 = hprev = hp;
      d0 = dprev = dp;
      run = 1;
    }
    return flush();
  };
  bool any_early = false;
  auto early = [&](uint32_t f) { return (frames[f].flags & VX_FRAME_INSERT_OCCLUSION_LABELS) == 0; };
  auto late_labels = [&](uint32_t f) { return (frames[f].flags & VX_FRAME_INSERT_OCCLUSION_LABELS) != 0; };  // they change during the traversal
  auto every = [&](uint32_t) { return true; };
  for (uint32_t g = 0; g < n_sub; ++g) {
    const uint32_t f0 = g * c->table_sets, f1 = f0 + c->table_sets < nb ? f0 + c->table_sets : nb;
    bool any = false;
    for (uint32_t f = f0; f < f1 && !any; ++f) any = early(f) && (frames[f].out_label || frames[f].out_bin);
    if (!any) continue;
    any_early = true;
    uint32_t n_copied = 0;
    VX_CUDA(c, cudaStreamWaitEvent(c->copy_out, ln.ev_sub[4 * g + 3], 0));
    VX_CUDA(c, copy_out_kind(f0, f1, OUT_LABEL, c->copy_out, early, &n_copied));
    VX_CUDA(c, copy_out_kind(f0, f1, OUT_BIN, c->copy_out, early, &n_copied));
  }
  {
    uint32_t n_copied = 0;
    VX_CUDA(c, copy_out_kind(0, nb, OUT_LABEL, st, late_labels, &n_copied));
    VX_CUDA(c, copy_out_kind(0, nb, OUT_BIN, st, late_labels, &n_copied));
    VX_CUDA(c, copy_out_kind(0, nb, OUT_OCCLUDED, st, every, &n_copied));
    VX_CUDA(c, copy_out_kind(0, nb, OUT_INVALID, st, every, &n_copied));
  }
  if (any_early) {
    VX_CUDA(c, cudaEventRecord(c->ev_copy_out, c->copy_out));
    VX_CUDA(c, cudaStreamWaitEvent(st, c->ev_copy_out, 0));
  }
  VX_CUDA(c, cudaMemcpyAsync(c->h_counters + (size_t)base * kCounterWords, c->d_counters + (size_t)base * kCounterWords,
                             (size_t)nb * kCounterWords * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
  VX_CUDA(c, cudaEventRecord(ln.ev[EV_END], st));
  ln.busy = true;
  ln.seq = ++c->batch_seq;
  trace_line("enqueue_end", ln.seq, nb);
  ln.frames = frames;
  ln.nb = nb;
  ln.n_sub = n_sub;
  ln.points = points;
  ln.fill_launches = fill_launches;
  return VX_OK;
}

// Waits for the batch on a lane and books its counters and times.  VX_ERR_TABLE_OVERFLOW: a histogram table was too small (the
// caller grows the tables and runs the batch again: ln.frames / ln.nb stay).  The lane stays `busy` until settle_batch.
int finish_batch(vx_ctx* c, Lane& ln) {
  const uint32_t base = ln.slot_base, nb = ln.nb, n_sub = ln.n_sub;
  const vx_frame_desc* frames = ln.frames;
  const uint64_t points = ln.points, fill_launches = ln.fill_launches;
  NvtxRange nvtx_wait("vx:wait");
  VX_CUDA(c, cudaStreamSynchronize(ln.stream));
  if (std::getenv("VX_B200_TRACE")) {  // stage boundaries of the batch, ms after its EV_BEGIN
    float a = 0, b = 0, v0 = 0, v1 = 0, e1 = 0;
    if (n_sub) {
      cudaEventElapsedTime(&a, ln.ev[EV_BEGIN], ln.ev_sub[0]);
      cudaEventElapsedTime(&b, ln.ev[EV_BEGIN], ln.ev_sub[4 * (n_sub - 1) + 3]);
    }
    cudaEventElapsedTime(&v0, ln.ev[EV_BEGIN], ln.ev[EV_VIS_BEGIN]);
    cudaEventElapsedTime(&v1, ln.ev[EV_BEGIN], ln.ev[EV_VIS]);
    cudaEventElapsedTime(&e1, ln.ev[EV_BEGIN], ln.ev[EV_END]);
    char buf[160];
    std::snprintf(buf, sizeof(buf), "first_fill %.1f last_emit %.1f vis_begin %.1f vis_end %.1f end %.1f", a, b, v0, v1, e1);
    trace_line("retire", ln.seq, nb, buf);
  }
  if (c->keep_tables) {
    c->tables_dirty = true;
    c->dirty_slots = nb;
  }
  c->last_batch = nb;
  c->last_base = base;

  bool overflow = false;
  unsigned long long stats[VX_SLOT_STATS] = {0};
  for (uint32_t f = 0; f < nb; ++f) {
    const uint32_t* w = c->h_counters + (size_t)(base + f) * kCounterWords;
    if (w[1] || w[3]) overflow = true;
    unsigned long long rs[VX_SLOT_STATS];
    std::memcpy(rs, w + 4, sizeof(rs));
    for (int k = 0; k < VX_SLOT_STATS; ++k) stats[k] += rs[k];
  }
  if (overflow) return VX_ERR_TABLE_OVERFLOW;
  // developer statistics, one line per frame: index, passes - 1, rays, steps, waves, resolve waves, sweeps, cycles of
  // the four phases, SM the frame ran on, start / end (globaltimer ns), free-and-occluded cells after pass 0 |
  // occupied cells << 32
  if (const char* path = std::getenv("VX_B200_FRAME_STATS")) {
    if (FILE* fp = std::fopen(path, "a")) {
      std::fprintf(fp, "# batch %llu\n", (unsigned long long)ln.seq);
      for (uint32_t f = 0; f < nb; ++f) {
        unsigned long long rs[VX_SLOT_STATS];
        std::memcpy(rs, c->h_counters + (size_t)(base + f) * kCounterWords + 4, sizeof(rs));
        std::fprintf(fp, "%u %u", f, frames[f].n_endpoints);
        for (int k = 0; k < VX_SLOT_STATS; ++k) std::fprintf(fp, " %llu", rs[k]);
        std::fprintf(fp, "\n");
      }
      std::fclose(fp);
    }
  }

  float ms = 0;
  auto span = [&](int a, int b) {
    cudaEventElapsedTime(&ms, ln.ev[a], ln.ev[b]);
    return (double)ms;
  };
  for (uint32_t g = 0; g < n_sub; ++g) {
    auto sub = [&](int a, int b) {
      cudaEventElapsedTime(&ms, ln.ev_sub[4 * g + a], ln.ev_sub[4 * g + b]);
      return (double)ms;
    };
    c->times.fill_ms += sub(0, 1);
    c->times.vote_ms += sub(1, 2);
    c->times.emit_ms += sub(2, 3);
  }
  c->times.visibility_ms += span(EV_VIS_BEGIN, EV_VIS);
  c->times.pack_ms += span(EV_VIS, EV_PACK);
  c->times.total_ms += span(EV_BEGIN, EV_END);
  c->times.frames += nb;
  c->times.points_read += points;
  c->times.fill_launches += fill_launches;
  c->times.vote_launches += 2 * n_sub;  // clear + vote
  c->times.emit_launches += n_sub;
  c->times.visibility_launches += 1;
  c->times.pack_launches += 1;
  c->times.rays_cast += stats[0];
  c->times.dda_steps += stats[1];
  c->times.vis_waves += stats[2];
  c->times.vis_resolve_waves += stats[3];
  c->times.vis_rounds += stats[4];
  for (int k = 0; k < 4; ++k) c->times.vis_cycles[k] += stats[5 + k];
  for (int k = 0; k < 3; ++k) c->times.vis_cycles[4 + k] += stats[13 + k];  // developer statistics (VX_TRACE_STATS builds)
  return VX_OK;
}

void evict_locked(vx_ctx* c, uint32_t scan_id) {
  const auto it = c->scans.find(scan_id);
  if (it == c->scans.end()) return;
  if (it->second.points) cudaFreeAsync(it->second.points, c->copy_in);
  if (it->second.labels) cudaFreeAsync(it->second.labels, c->copy_in);
  c->scans.erase(it);
}

// ---- the worker: jobs in submission order, their batches alternating between the lanes -------------------------------------

Lane* oldest_busy(vx_ctx* c) {
  Lane* best = nullptr;
  for (uint32_t k = 0; k < kMaxLanes; ++k)
    if (c->lanes[k].busy && (!best || c->lanes[k].seq < best->seq)) best = &c->lanes[k];
  return best;
}

Lane* newest_busy(vx_ctx* c) {
  Lane* best = nullptr;
  for (uint32_t k = 0; k < kMaxLanes; ++k)
    if (c->lanes[k].busy && (!best || c->lanes[k].seq > best->seq)) best = &c->lanes[k];
  return best;
}

// The batch on `ln` is over (rc = how): book it on its job; the job's last batch publishes the job.
void settle_batch(vx_ctx* c, Lane& ln, int rc, const std::string& error) {
  Job* job = ln.job;
  if (rc != VX_OK && job->rc == VX_OK) {
    job->rc = rc;
    job->error = error;
  }
  ln.busy = false;
  ln.frames = nullptr;
  ln.job = nullptr;
  if (ln.owned) {
    std::unique_ptr<Job> done = std::move(ln.owned);
    {
      std::lock_guard<std::mutex> lk(c->qmu);
      if (done->rc != VX_OK) c->failed[done->ticket] = std::make_pair(done->rc, done->error);
      c->finished_ticket = done->ticket;
    }
    c->dcv.notify_all();
  }
}

int grow_tables(vx_ctx* c) {  // x4; every lane is idle
  if (c->log2_target >= 28) return fail(c, VX_ERR_TABLE_OVERFLOW, "histogram table overflow at 2^28 slots");
  for (uint32_t k = 0; k < kMaxLanes; ++k) {
    if (c->lanes[k].stream) VX_CUDA(c, cudaStreamSynchronize(c->lanes[k].stream));
    if (c->lanes[k].fill_stream) VX_CUDA(c, cudaStreamSynchronize(c->lanes[k].fill_stream));
  }
  free_slots(c);
  c->log2_target += 2;
  c->log2_input = c->log2_target > 12 ? c->log2_target - 3 : c->log2_target;
  size_t free_b = 0, total_b = 0;
  VX_CUDA(c, cudaMemGetInfo(&free_b, &total_b));
  while (c->table_sets > 1 && (size_t)c->table_sets * table_set_bytes(c) > free_b / 2) c->table_sets /= 2;
  return VX_OK;
}

// Completes the oldest batch in flight.  If its tables overflowed, every other lane is completed too (the tables are rebuilt and
// the frame records go with them), the tables grow and the batches run again, one at a time and oldest first; then everything is
// settled, oldest first.
void retire_oldest(vx_ctx* c) {
  Lane* ln = oldest_busy(c);
  if (!ln) return;
  std::string err;
  tl_error_sink = &err;
  int rc = finish_batch(c, *ln);
  if (rc != VX_ERR_TABLE_OVERFLOW) {
    tl_error_sink = nullptr;
    settle_batch(c, *ln, rc, err);
    return;
  }
  struct Redo {
    Lane* lane;
    int rc;
    std::string err;
  };
  std::vector<Redo> newer;  // the other batches in flight, oldest first
  for (uint32_t k = 0; k < kMaxLanes; ++k)
    if (c->lanes[k].busy && &c->lanes[k] != ln) newer.push_back(Redo{&c->lanes[k], VX_OK, std::string()});
  std::sort(newer.begin(), newer.end(), [](const Redo& a, const Redo& b) { return a.lane->seq < b.lane->seq; });
  for (Redo& r : newer) {
    tl_error_sink = &r.err;
    r.rc = finish_batch(c, *r.lane);
  }
  auto rerun = [&](Lane& l, std::string& e, bool grow_first) {
    tl_error_sink = &e;
    for (bool grow = grow_first;; grow = true) {
      int r = grow ? grow_tables(c) : VX_OK;
      if (r != VX_OK) return r;
      // (every lane is idle here -- finish_batch waited for each -- and nothing of another batch is waited for: prev = null)
      r = enqueue_batch(c, l, nullptr, l.frames, l.nb);
      if (r != VX_OK) return r;
      r = finish_batch(c, l);
      if (r != VX_ERR_TABLE_OVERFLOW) return r;
    }
  };
  rc = rerun(*ln, err, true);
  for (Redo& r : newer) {
    // the frame records of the newer batches went with the old tables: run them again whatever they said (a CUDA error stays)
    if (r.rc == VX_ERR_TABLE_OVERFLOW || r.rc == VX_OK) {
      r.err.clear();
      r.rc = rerun(*r.lane, r.err, false);
    }
  }
  tl_error_sink = nullptr;
  settle_batch(c, *ln, rc, err);
  for (Redo& r : newer) settle_batch(c, *r.lane, r.rc, r.err);
}

// Enqueues every batch of a job; its last batch stays in flight (the job is published when that batch is settled).
void run_job(vx_ctx* c, std::unique_ptr<Job> job) {
  Job* j = job.get();
  const uint32_t n = (uint32_t)j->frames.size();
  auto abandon = [&](int rc, const std::string& error) {  // nothing of this job is in flight any more: publish the failure in order
    while (oldest_busy(c)) retire_oldest(c);
    if (j->rc == VX_OK) {
      j->rc = rc;
      j->error = error;
    }
    {
      std::lock_guard<std::mutex> lk(c->qmu);
      if (j->rc != VX_OK) c->failed[j->ticket] = std::make_pair(j->rc, j->error);
      c->finished_ticket = j->ticket;
    }
    c->dcv.notify_all();
  };
  if (cudaSetDevice(c->device) != cudaSuccess) return abandon(VX_ERR_CUDA, "cudaSetDevice failed");
  if (n == 0) return abandon(VX_OK, "");
  // batches of equal size (a short last batch would leave most of the device idle for a whole batch time)
  const uint32_t n_batches = (n + c->lane_cap - 1) / c->lane_cap;
  const uint32_t per_batch = (n + n_batches - 1) / n_batches;
  for (uint32_t base = 0; base < n;) {
    const uint32_t nb = n - base < per_batch ? n - base : per_batch;
    Lane& ln = c->lanes[c->next_lane];
    while (ln.busy) retire_oldest(c);
    if (j->rc != VX_OK) return abandon(j->rc, j->error);  // an earlier batch of this job failed
    std::string err;
    tl_error_sink = &err;
    const int rc = enqueue_batch(c, ln, newest_busy(c), j->frames.data() + base, nb);
    tl_error_sink = nullptr;
    if (rc != VX_OK) {
      if (ln.fill_stream) cudaStreamSynchronize(ln.fill_stream);
      cudaStreamSynchronize(ln.stream);
      return abandon(rc, err);
    }
    ln.job = j;
    base += nb;
    if (base >= n) ln.owned = std::move(job);
    c->next_lane = (c->next_lane + 1u) % c->depth;
  }
}

void worker_main(vx_ctx* c) {
  for (;;) {
    std::unique_ptr<Job> job;
    Lane* oldest = oldest_busy(c);
    {
      std::unique_lock<std::mutex> lk(c->qmu);
      if (!oldest) {
        c->qcv.wait(lk, [&] { return c->stop || !c->queue.empty(); });
      } else if (c->queue.empty()) {
        // A batch is in flight and nothing is queued: do not block in the driver -- a job submitted meanwhile must be enqueued
        // behind it at once (that is the overlap) -- but look at the batch's end event every 100 us.
        if (cudaEventQuery(oldest->ev[EV_END]) != cudaSuccess) {
          cudaGetLastError();
          c->qcv.wait_for(lk, std::chrono::microseconds(100), [&] { return !c->queue.empty(); });
        }
      }
      if (!c->queue.empty()) {
        job = std::move(c->queue.front());
        c->queue.pop_front();
      } else if (!oldest) {
        return;  // stop requested, nothing queued, nothing in flight
      }
    }
    if (job) {
      run_job(c, std::move(job));
    } else if (cudaEventQuery(oldest->ev[EV_END]) == cudaSuccess) {
      retire_oldest(c);
    } else {
      cudaGetLastError();
    }
  }
}

}  // namespace

extern "C" {

const char* vx_last_global_error(void) { return g_global_error.c_str(); }
const char* vx_last_error(const vx_ctx* ctx) { return ctx ? ctx->error.c_str() : g_global_error.c_str(); }

int vx_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
  return n;
}

void* vx_host_alloc(size_t bytes) {
  void* p = nullptr;
  if (cudaMallocHost(&p, bytes) != cudaSuccess) return nullptr;
  return p;
}

void vx_host_free(void* p) {
  if (p) cudaFreeHost(p);
}

int vx_create(int device, const vx_grid_desc* grid, const vx_filter_desc* filter, const vx_options* options, vx_ctx** out) {
  if (!grid || !filter || !out) return fail(nullptr, VX_ERR_INVALID, "null argument");
  *out = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0)
    return fail(nullptr, VX_ERR_NO_DEVICE, "no CUDA device: voxelizer_b200 has no CPU path");
  if (device < 0 || device >= ndev) return fail(nullptr, VX_ERR_INVALID, "device index out of range");
  if (!(grid->resolution > 0.0f)) return fail(nullptr, VX_ERR_INVALID, "resolution must be positive");

  vx_ctx* c = new vx_ctx;
  auto bail = [&](int code) {
    g_global_error = c->error;
    vx_destroy(c);
    return code;
  };
#define VX_CREATE_CUDA(call)                                                          \
  do {                                                                                \
    cudaError_t e__ = (call);                                                         \
    if (e__ != cudaSuccess) {                                                         \
      c->error = std::string(#call) + ": " + cudaGetErrorString(e__);                 \
      return bail(VX_ERR_CUDA);                                                       \
    }                                                                                 \
  } while (0)

  c->device = device;
  VX_CREATE_CUDA(cudaSetDevice(device));
  if (options && options->stream) {
    c->stream = static_cast<cudaStream_t>(options->stream);
  } else {
    VX_CREATE_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    c->own_stream = true;
  }
  c->lanes[0].stream = c->stream;
  {
    // developer switches (A/B on the device, profiles/r02_lanes_shapes_ab.txt): VX_B200_LANES = batches in flight (1 .. 4,
    // default 2), VX_B200_FILL_PRIO=1 puts the fill / vote / emit kernels on a high-priority stream of the lane
    int lo_prio = 0, hi_prio = 0;
    VX_CREATE_CUDA(cudaDeviceGetStreamPriorityRange(&lo_prio, &hi_prio));
    const char* e = std::getenv("VX_B200_FILL_PRIO");
    const bool fill_prio = e && std::atoi(e) != 0 && hi_prio < lo_prio;
    for (uint32_t k = 0; k < kMaxLanes; ++k) {
      Lane& ln = c->lanes[k];
      if (k > 0) VX_CREATE_CUDA(cudaStreamCreateWithFlags(&ln.stream, cudaStreamNonBlocking));
      for (int i = 0; i < EV_COUNT; ++i) VX_CREATE_CUDA(cudaEventCreate(&ln.ev[i]));
      if (fill_prio) {
        VX_CREATE_CUDA(cudaStreamCreateWithPriority(&ln.fill_stream, cudaStreamNonBlocking, hi_prio));
        VX_CREATE_CUDA(cudaEventCreateWithFlags(&ln.ev_staged, cudaEventDisableTiming));
        VX_CREATE_CUDA(cudaEventCreateWithFlags(&ln.ev_filled, cudaEventDisableTiming));
      }
    }
  }
  VX_CREATE_CUDA(cudaStreamCreateWithFlags(&c->copy_in, cudaStreamNonBlocking));
  VX_CREATE_CUDA(cudaStreamCreateWithFlags(&c->copy_out, cudaStreamNonBlocking));
  VX_CREATE_CUDA(cudaEventCreateWithFlags(&c->ev_copy_out, cudaEventDisableTiming));
  {  // scans live in a stream-ordered memory pool OF THIS CONTEXT (the device's default pool belongs to the embedding
     // application); freed blocks are kept for the next upload and returned to the driver by vx_destroy
    cudaMemPoolProps props{};
    props.allocType = cudaMemAllocationTypePinned;
    props.handleTypes = cudaMemHandleTypeNone;
    props.location.type = cudaMemLocationTypeDevice;
    props.location.id = device;
    VX_CREATE_CUDA(cudaMemPoolCreate(&c->scan_pool, &props));
    uint64_t keep = UINT64_MAX;
    VX_CREATE_CUDA(cudaMemPoolSetAttribute(c->scan_pool, cudaMemPoolAttrReleaseThreshold, &keep));
  }
  VX_CREATE_CUDA(vx_visibility_prepare());

  derive_geometry(*grid, c->geom);
  std::memcpy(c->min_extent, grid->min_extent, sizeof(c->min_extent));
  std::memcpy(c->max_extent, grid->max_extent, sizeof(c->max_extent));
  if (c->geom.nx <= 0 || c->geom.ny <= 0 || c->geom.nz <= 0) {
    c->error = "empty grid";
    return bail(VX_ERR_INVALID);
  }
  c->nvox = (uint64_t)c->geom.nx * (uint64_t)c->geom.ny * (uint64_t)c->geom.nz;
  if (c->nvox >= (1ull << 31) || vx_visibility_visits(c->geom) >= (1ull << 25) || c->geom.nx > 2047 || c->geom.ny > 2047 ||
      c->geom.nz > 1023) {
    c->error = "grid too large: the traversal packs positions into 11 + 11 + 10 bits and counts visits in 25 bits";
    return bail(VX_ERR_UNSUPPORTED);
  }
  c->words = (uint32_t)((c->nvox + 31) / 32);
  c->visits = vx_visibility_visits(c->geom);

  // label filter LUT (voxelize_utils.cpp:175-180,195-198): label -> joined label -> in `ignore`?
  c->filter.min_range = filter->min_range;
  c->filter.max_range = filter->max_range;
  c->filter.hidecar = filter->hidecar ? 1 : 0;
  c->filter.has_label_filter = filter->n_filtered > 0 ? 1 : 0;
  {
    std::map<uint32_t, uint32_t> member_to_key;  // later keys overwrite earlier ones, as in the reference
    for (uint32_t i = 0; i < filter->n_join_pairs; ++i) member_to_key[filter->join_pairs[2 * i + 1]] = filter->join_pairs[2 * i];
    std::vector<uint32_t> bits(65536 / 32, 0);
    for (uint32_t label = 0; label < 65536; ++label) {
      uint32_t mapped = label;
      const auto it = member_to_key.find(label);
      if (it != member_to_key.end()) mapped = it->second;
      bool rejected = false;
      for (uint32_t i = 0; i < filter->n_filtered; ++i) rejected |= filter->filtered_labels[i] == mapped;
      if (rejected) bits[label >> 5] |= 1u << (label & 31);
    }
    VX_CREATE_CUDA(cudaMalloc(reinterpret_cast<void**>(&c->d_reject_bits), bits.size() * 4));
    VX_CREATE_CUDA(cudaMemcpy(c->d_reject_bits, bits.data(), bits.size() * 4, cudaMemcpyHostToDevice));
    c->filter.reject_bits = c->d_reject_bits;
  }

  if (options) {
    if (options->table_log2_capacity) c->log2_target = options->table_log2_capacity;
    c->keep_tables = (options->debug & 1u) != 0;
    c->vis_mode = options->vis_mode <= 3u ? (int)options->vis_mode : 0;
    c->occlusion_labels = options->occlusion_labels != 0;
  }
  if (c->log2_target < 10) c->log2_target = 10;
  if (c->log2_target > 30) c->log2_target = 30;
  c->log2_input = c->log2_target > 12 ? c->log2_target - 3 : c->log2_target;

  // Sizing.  A batch = the frames of ONE visibility launch (one CTA per frame, scheduled by the hardware
  // as resident CTAs retire); its histogram tables are shared by frames `table_sets` apart.
  const uint32_t resident = (uint32_t)vx_visibility_resident_ctas(device);
  if (resident == 0) {
    c->error = "the visibility kernel does not fit on this device";
    return bail(VX_ERR_CUDA);
  }
  c->table_sets = options && options->table_sets ? options->table_sets : 128u;
  size_t free_b = 0, total_b = 0;
  VX_CREATE_CUDA(cudaMemGetInfo(&free_b, &total_b));
  uint32_t want = options && options->frames_in_flight ? options->frames_in_flight : 4u * resident;
  if (c->keep_tables && want > 8u && !(options && options->frames_in_flight)) want = 8u;  // debug: small batches,
  if (c->keep_tables || want < c->table_sets) c->table_sets = want;                       // every frame its own tables
  // Two batches in flight (two lanes, each with frame records for a whole batch) unless the tables are kept for debugging or a
  // scratch buffer per resident CTA (needed once CTAs of several launches share the SMs) does not fit next to everything else.
  // (Measured: with two lanes the 1036 traversal slots are already full all the time; three or four lanes give the same
  // 495 frames/s, with or without the high-priority fill stream.)
  if (const char* e = std::getenv("VX_B200_LEAVE_ROOM")) c->leave_room = std::atoi(e) != 0;
  c->depth = 2;
  if (const char* e = std::getenv("VX_B200_LANES")) {
    const int n = std::atoi(e);
    if (n >= 1 && n <= (int)kMaxLanes) c->depth = (uint32_t)n;
  }
  if (c->keep_tables || (size_t)resident * scratch_bytes(c) > free_b / 4) c->depth = 1;
  {  // the largest batch whose frame records, histogram tables and traversal scratch fit into 70 % of the free memory
     // (scratch: one buffer per RESIDENT CTA -- at most one per frame with one lane; tables: one set per frame of a fill group)
    const size_t budget = free_b / 10 * 7;
    auto cost = [&](uint32_t n) {
      const size_t n_scratch = (c->depth == 1 && n < resident) ? n : resident;
      const size_t n_sets = n < c->table_sets ? n : c->table_sets;
      return n_scratch * scratch_bytes(c) + n_sets * table_set_bytes(c) + (size_t)c->depth * n * frame_bytes(c);
    };
    while (want > 1 && cost(want) > budget) want = want - (want + 3) / 4;
  }
  if (want > 32767u) want = 32767u;
  if (want < 1) want = 1;
  if (c->table_sets > want) c->table_sets = want;
  c->lane_cap = want;                 // frames per batch
  c->max_slots = c->depth * want;     // frame records (allocated as batches need them)
  for (uint32_t k = 0; k < kMaxLanes; ++k) c->lanes[k].slot_base = k * c->lane_cap;  // (lanes >= depth are never used)
  VX_CREATE_CUDA(cudaMalloc(reinterpret_cast<void**>(&c->d_slots), (size_t)c->max_slots * sizeof(VxSlot)));
  VX_CREATE_CUDA(cudaMalloc(reinterpret_cast<void**>(&c->d_counters), (size_t)c->max_slots * kCounterWords * sizeof(uint32_t)));
  VX_CREATE_CUDA(cudaMallocHost(reinterpret_cast<void**>(&c->h_counters), (size_t)c->max_slots * kCounterWords * sizeof(uint32_t)));
#undef VX_CREATE_CUDA
  *out = c;
  return VX_OK;
}

void vx_destroy(vx_ctx* c) {
  if (!c) return;
  if (c->worker.joinable()) {  // pending jobs are finished first
    {
      std::lock_guard<std::mutex> lk(c->qmu);
      c->stop = true;
    }
    c->qcv.notify_all();
    c->worker.join();
  }
  cudaSetDevice(c->device);
  for (uint32_t k = 0; k < kMaxLanes; ++k) {
    if (c->lanes[k].fill_stream) cudaStreamSynchronize(c->lanes[k].fill_stream);
    if (c->lanes[k].stream) cudaStreamSynchronize(c->lanes[k].stream);
  }
  if (c->copy_out) cudaStreamSynchronize(c->copy_out);
  if (c->copy_in) cudaStreamSynchronize(c->copy_in);
  free_slots(c);
  for (auto& kv : c->scans) {
    if (kv.second.points) cudaFreeAsync(kv.second.points, c->copy_in);
    if (kv.second.labels) cudaFreeAsync(kv.second.labels, c->copy_in);
  }
  if (c->copy_in) cudaStreamSynchronize(c->copy_in);
  if (c->scan_pool) cudaMemPoolDestroy(c->scan_pool);
  for (const UploadMark& m : c->marks) cudaEventDestroy(m.ev);
  for (cudaEvent_t e : c->ev_pool) cudaEventDestroy(e);
  if (c->ev_copy_out) cudaEventDestroy(c->ev_copy_out);
  if (c->copy_out) cudaStreamDestroy(c->copy_out);
  if (c->copy_in) cudaStreamDestroy(c->copy_in);
  for (void* p : c->scratch_slabs) cudaFree(p);
  if (c->d_scratch) cudaFree(c->d_scratch);
  if (c->d_locks) cudaFree(c->d_locks);
  for (uint32_t k = 0; k < kMaxLanes; ++k) {
    Lane& ln = c->lanes[k];
    for (cudaEvent_t e : ln.ev_sub) cudaEventDestroy(e);
    for (int i = 0; i < EV_COUNT; ++i)
      if (ln.ev[i]) cudaEventDestroy(ln.ev[i]);
    if (ln.ev_staged) cudaEventDestroy(ln.ev_staged);
    if (ln.ev_filled) cudaEventDestroy(ln.ev_filled);
    if (ln.h_stage) cudaFreeHost(ln.h_stage);
    if (ln.d_stage) cudaFree(ln.d_stage);
    if (ln.fill_stream) cudaStreamDestroy(ln.fill_stream);
    if (k > 0 && ln.stream) cudaStreamDestroy(ln.stream);
  }
  if (c->d_slots) cudaFree(c->d_slots);
  if (c->d_counters) cudaFree(c->d_counters);
  if (c->h_counters) cudaFreeHost(c->h_counters);
  if (c->d_reject_bits) cudaFree(c->d_reject_bits);
  if (c->stream && c->own_stream) cudaStreamDestroy(c->stream);
  delete c;
}

int vx_grid_info_get(const vx_ctx* c, vx_grid_info* out) {
  if (!c || !out) return VX_ERR_INVALID;
  out->size[0] = (uint32_t)c->geom.nx;
  out->size[1] = (uint32_t)c->geom.ny;
  out->size[2] = (uint32_t)c->geom.nz;
  out->offset[0] = c->geom.off[0];
  out->offset[1] = c->geom.off[1];
  out->offset[2] = c->geom.off[2];
  out->resolution = c->geom.res;
  out->num_voxels = c->nvox;
  out->packed_bytes = c->nvox / 8;
  out->label_bytes = c->nvox * 2;
  return VX_OK;
}

uint32_t vx_frames_in_flight(const vx_ctx* c) { return c ? c->lane_cap : 0; }
uint32_t vx_batches_in_flight(const vx_ctx* c) { return c ? c->depth : 0; }

int vx_upload_scan(vx_ctx* c, uint32_t scan_id, const float* xyzr, const uint32_t* labels, uint32_t n) {
  if (!c || (n && (!xyzr || !labels))) return fail(c, VX_ERR_INVALID, "null argument");
  VX_CUDA(c, cudaSetDevice(c->device));
  std::lock_guard<std::mutex> lk(c->mu);
  evict_locked(c, scan_id);
  DeviceScan s;
  s.n = n;
  if (n) {
    // stream-ordered pool allocation on the upload stream: no device-wide synchronisation per scan.  A block freed
    // by an evict is handed out again at once: the caller must not evict a scan that a submitted, not yet awaited
    // job reads (include/voxelizer_b200.h).
    VX_CUDA(c, cudaMallocFromPoolAsync(reinterpret_cast<void**>(&s.points), (size_t)n * sizeof(float4), c->scan_pool, c->copy_in));
    VX_CUDA(c, cudaMallocFromPoolAsync(reinterpret_cast<void**>(&s.labels), (size_t)n * sizeof(uint32_t), c->scan_pool, c->copy_in));
    VX_CUDA(c, cudaMemcpyAsync(s.points, xyzr, (size_t)n * sizeof(float4), cudaMemcpyHostToDevice, c->copy_in));
    VX_CUDA(c, cudaMemcpyAsync(s.labels, labels, (size_t)n * sizeof(uint32_t), cudaMemcpyHostToDevice, c->copy_in));
    s.seq = ++c->upload_seq;
    if (c->upload_seq - c->marked_seq >= kUploadsPerMark) {
      const int rc = mark_uploads(c);
      if (rc != VX_OK) return rc;
    }
  }
  c->scans[scan_id] = s;
  return VX_OK;
}

uint64_t vx_upload_ticket(const vx_ctx* c) {
  if (!c) return 0;
  std::lock_guard<std::mutex> lk(const_cast<vx_ctx*>(c)->mu);
  return c->upload_seq;
}

int vx_upload_wait(vx_ctx* c, uint64_t ticket) {
  if (!c) return VX_ERR_INVALID;
  cudaEvent_t ev = nullptr;
  {
    std::lock_guard<std::mutex> lk(c->mu);
    if (ticket > c->upload_seq) ticket = c->upload_seq;
    if (ticket <= c->done_seq) return VX_OK;
    VX_CUDA(c, cudaSetDevice(c->device));
    if (ticket > c->marked_seq) {
      const int rc = mark_uploads(c);
      if (rc != VX_OK) return rc;
    }
    for (const UploadMark& m : c->marks)
      if (m.seq >= ticket) {
        ev = m.ev;
        break;
      }
  }
  if (ev) VX_CUDA(c, cudaEventSynchronize(ev));  // (unlocked: a job may be enqueueing meanwhile)
  std::lock_guard<std::mutex> lk(c->mu);
  prune_marks(c);
  return VX_OK;
}

int vx_evict_scan(vx_ctx* c, uint32_t scan_id) {
  if (!c) return VX_ERR_INVALID;
  std::lock_guard<std::mutex> lk(c->mu);
  evict_locked(c, scan_id);
  return VX_OK;
}

int vx_evict_all(vx_ctx* c) {
  if (!c) return VX_ERR_INVALID;
  std::lock_guard<std::mutex> lk(c->mu);
  for (auto& kv : c->scans) {
    if (kv.second.points) cudaFreeAsync(kv.second.points, c->copy_in);
    if (kv.second.labels) cudaFreeAsync(kv.second.labels, c->copy_in);
  }
  c->scans.clear();
  return VX_OK;
}

int vx_submit_frames(vx_ctx* c, const vx_frame_desc* frames, uint32_t n, uint64_t* ticket) {
  if (!c || (n && !frames) || !ticket) return fail(c, VX_ERR_INVALID, "null argument");
  std::unique_ptr<Job> job(new Job);
  job->frames.assign(frames, frames + n);
  size_t n_ids = 0, n_ap = 0, n_ep = 0;
  for (uint32_t f = 0; f < n; ++f) {
    const vx_frame_desc& fr = frames[f];
    if ((fr.n_prior && (!fr.prior_ids || !fr.ap_prior)) || (fr.n_past && (!fr.past_ids || !fr.ap_past)) || (fr.n_endpoints && !fr.endpoints))
      return fail(c, VX_ERR_INVALID, "frame descriptor has null arrays");
    n_ids += fr.n_prior + fr.n_past;
    n_ap += fr.n_prior + fr.n_past;
    n_ep += fr.n_endpoints;
  }
  job->ids.resize(n_ids);
  job->ap.resize(n_ap * 16);
  job->ep.resize(n_ep * 3);
  size_t ii = 0, ai = 0, ei = 0;
  for (uint32_t f = 0; f < n; ++f) {  // (the vectors never grow again: the pointers below stay valid)
    vx_frame_desc& fr = job->frames[f];
    auto take_ids = [&](const uint32_t*& p, uint32_t k) {
      if (k) std::memcpy(&job->ids[ii], p, (size_t)k * 4);
      p = k ? &job->ids[ii] : nullptr;
      ii += k;
    };
    auto take_ap = [&](const float*& p, uint32_t k) {
      if (k) std::memcpy(&job->ap[ai * 16], p, (size_t)k * 64);
      p = k ? &job->ap[ai * 16] : nullptr;
      ai += k;
    };
    take_ids(fr.prior_ids, fr.n_prior);
    take_ap(fr.ap_prior, fr.n_prior);
    take_ids(fr.past_ids, fr.n_past);
    take_ap(fr.ap_past, fr.n_past);
    if (fr.n_endpoints) std::memcpy(&job->ep[ei * 3], fr.endpoints, (size_t)fr.n_endpoints * 12);
    fr.endpoints = fr.n_endpoints ? &job->ep[ei * 3] : nullptr;
    ei += fr.n_endpoints;
  }
  {
    std::lock_guard<std::mutex> lk(c->qmu);
    if (!c->worker.joinable()) c->worker = std::thread(worker_main, c);
    job->ticket = c->next_ticket++;
    *ticket = job->ticket;
    trace_line("submit", job->ticket, n);
    c->queue.push_back(std::move(job));
  }
  c->qcv.notify_one();
  return VX_OK;
}

int vx_wait(vx_ctx* c, uint64_t ticket) {
  if (!c) return VX_ERR_INVALID;
  std::unique_lock<std::mutex> lk(c->qmu);
  if (ticket == 0 || ticket >= c->next_ticket) ticket = c->next_ticket - 1;  // 0 = everything submitted so far
  c->dcv.wait(lk, [&] { return c->finished_ticket >= ticket; });
  int rc = VX_OK;
  for (auto it = c->failed.begin(); it != c->failed.end() && it->first <= ticket;) {  // the first failure up to `ticket`
    if (rc == VX_OK) {
      rc = it->second.first;
      c->error = it->second.second;
    }
    it = c->failed.erase(it);
  }
  return rc;
}

int vx_process_frames(vx_ctx* c, const vx_frame_desc* frames, uint32_t n) {
  if (!c || (n && !frames)) return fail(c, VX_ERR_INVALID, "null argument");
  uint64_t ticket = 0;
  const int rc = vx_submit_frames(c, frames, n, &ticket);
  if (rc != VX_OK) return rc;
  return vx_wait(c, ticket);
}

int vx_sync(vx_ctx* c) {
  if (!c) return VX_ERR_INVALID;
  const int rc = vx_wait(c, 0);
  VX_CUDA(c, cudaSetDevice(c->device));
  VX_CUDA(c, cudaStreamSynchronize(c->copy_in));
  for (uint32_t k = 0; k < kMaxLanes; ++k)
    if (c->lanes[k].stream) VX_CUDA(c, cudaStreamSynchronize(c->lanes[k].stream));
  std::lock_guard<std::mutex> lk(c->mu);
  prune_marks(c);
  return rc;
}

int vx_stage_times_get(vx_ctx* c, vx_stage_times* out) {
  if (!c || !out) return VX_ERR_INVALID;
  *out = c->times;
  return VX_OK;
}

int vx_stage_times_reset(vx_ctx* c) {
  if (!c) return VX_ERR_INVALID;
  c->times = vx_stage_times{};
  return VX_OK;
}

int vx_trace_ray(vx_ctx* c, uint32_t frame_index, int32_t i, int32_t j, int32_t k, const float* endpoint, int32_t* visited_ijk,
                 uint32_t visited_capacity, uint32_t* n_visited, int32_t* occluder) {
  if (!c || !endpoint || !occluder || (visited_capacity && !visited_ijk)) return fail(c, VX_ERR_INVALID, "null argument");
  if (frame_index >= c->last_batch) return fail(c, VX_ERR_INVALID, "frame index outside the last batch");
  VX_CUDA(c, cudaSetDevice(c->device));
  int32_t* d_buf = nullptr;
  const size_t words = 2 + (size_t)visited_capacity * 3;
  VX_CUDA(c, cudaMalloc(reinterpret_cast<void**>(&d_buf), words * 4));
  std::vector<int32_t> host(words);
  cudaError_t e = vx_launch_trace_ray(c->slots[c->last_base + frame_index].dev.occ, c->geom, i, j, k, endpoint, d_buf + 2, visited_capacity, d_buf, c->stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(host.data(), d_buf, words * 4, cudaMemcpyDeviceToHost, c->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
  cudaFree(d_buf);
  if (e != cudaSuccess) return fail(c, VX_ERR_CUDA, std::string("vx_trace_ray: ") + cudaGetErrorString(e));
  const uint32_t n = (uint32_t)host[0];
  if (n_visited) *n_visited = n;
  *occluder = host[1];
  const uint32_t keep = n < visited_capacity ? n : visited_capacity;
  if (keep) std::memcpy(visited_ijk, host.data() + 2, (size_t)keep * 12);
  return VX_OK;
}

int vx_debug_dump(vx_ctx* c, uint32_t frame_index, int kind, void* dst, uint64_t* count) {
  if (!c || !count) return fail(c, VX_ERR_INVALID, "null argument");
  if (frame_index >= c->last_batch) return fail(c, VX_ERR_INVALID, "frame index outside the last batch");
  VX_CUDA(c, cudaSetDevice(c->device));
  const VxSlot& s = c->slots[c->last_base + frame_index].dev;
  if (kind == VX_DUMP_HIST_TARGET || kind == VX_DUMP_HIST_INPUT) {
    if (!c->keep_tables || !c->tables_dirty) return fail(c, VX_ERR_INVALID, "histogram dump needs vx_options.debug = 1");
    const VxTable& t = kind == VX_DUMP_HIST_TARGET ? s.target : s.input;
    const size_t cap = (size_t)1 << t.log2_cap;
    std::vector<unsigned long long> keys(cap);
    std::vector<uint32_t> counts(cap);
    VX_CUDA(c, cudaMemcpy(keys.data(), t.keys, cap * 8, cudaMemcpyDeviceToHost));
    VX_CUDA(c, cudaMemcpy(counts.data(), t.counts, cap * 4, cudaMemcpyDeviceToHost));
    uint64_t used = 0;
    uint32_t* o = static_cast<uint32_t*>(dst);
    for (size_t h = 0; h < cap; ++h) {
      if (keys[h] == 0ull) continue;
      if (o) {
        const unsigned long long k = keys[h] - 1ull;
        o[3 * used + 0] = (uint32_t)(k >> 16);
        o[3 * used + 1] = (uint32_t)(k & 0xFFFFull);
        o[3 * used + 2] = counts[h];
      }
      ++used;
    }
    *count = used;
    return VX_OK;
  }
  if (kind == VX_DUMP_OCCUPANCY || kind == VX_DUMP_OCCLUDED || kind == VX_DUMP_INVALID) {
    *count = c->nvox;
    if (!dst) return VX_OK;
    std::vector<uint32_t> a(c->words), b(c->words);
    const uint32_t* src = kind == VX_DUMP_OCCUPANCY ? s.occ : (kind == VX_DUMP_OCCLUDED ? s.occl : s.alive);
    VX_CUDA(c, cudaMemcpy(a.data(), src, (size_t)c->words * 4, cudaMemcpyDeviceToHost));
    if (kind == VX_DUMP_INVALID) {
      VX_CUDA(c, cudaMemcpy(b.data(), s.occ, (size_t)c->words * 4, cudaMemcpyDeviceToHost));
      for (uint32_t w = 0; w < c->words; ++w) a[w] &= ~b[w];
    }
    uint8_t* o = static_cast<uint8_t*>(dst);
    for (uint64_t L = 0; L < c->nvox; ++L) o[L] = (a[L >> 5] >> (L & 31)) & 1u;
    return VX_OK;
  }
  if (kind == VX_DUMP_FILL_SOURCE) {
    *count = c->nvox;
    if (!dst) return VX_OK;
    if (!s.fill_source) return fail(c, VX_ERR_INVALID, "fill-source dump needs vx_options.occlusion_labels");
    VX_CUDA(c, cudaMemcpy(dst, s.fill_source, c->nvox * 4, cudaMemcpyDeviceToHost));
    return VX_OK;
  }
  return fail(c, VX_ERR_INVALID, "unknown dump kind");
}

}  // extern "C"
