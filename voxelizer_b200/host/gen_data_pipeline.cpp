#include "gen_data_pipeline.h"

#include <sys/stat.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <iostream>
#include <map>
#include <mutex>
#include <stdexcept>
#include <thread>

namespace vxb {

namespace {

double now() {
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

bool makePath(const std::string& path) {
  std::string cur;
  for (size_t i = 0; i <= path.size(); ++i) {
    if (i == path.size() || path[i] == '/') {
      if (!cur.empty()) {
        struct stat st;
        if (::stat(cur.c_str(), &st) != 0) {
          if (::mkdir(cur.c_str(), 0777) != 0) return false;
        } else if (!S_ISDIR(st.st_mode)) {
          return false;
        }
      }
    }
    if (i < path.size()) cur.push_back(path[i]);
  }
  return true;
}

bool writeAll(const std::string& name, const void* data, size_t bytes) {
  FILE* f = std::fopen(name.c_str(), "wb");
  if (!f) return false;
  const size_t n = std::fwrite(data, 1, bytes, f);
  return std::fclose(f) == 0 && n == bytes;
}

struct ResidentScan {
  Matrix4f pose;
  uint32_t points = 0;
  uint32_t labelled = 0;  // labels > 0 after the 0xFFFF mask
};

// One output set per frame of a batch, in PAGEABLE memory: pinning ~5 MB per frame costs seconds for a whole
// sequence (more than reading it), while a device -> pageable copy only blocks this thread, which has nothing else
// to do -- `.label` / `.bin` leave the device while the rays are traced (vx_api.cu: copy-out), the page faults of the
// first touch included.
struct OutputRing {
  uint8_t* base = nullptr;
  size_t perFrame = 0, labelBytes = 0, packedBytes = 0;
  uint32_t frames = 0;
  bool allocate(uint32_t n, const vx_grid_info& gi) {
    labelBytes = gi.label_bytes;
    packedBytes = gi.packed_bytes;
    perFrame = ((labelBytes + 3 * packedBytes + 255) / 256) * 256;
    base = static_cast<uint8_t*>(std::malloc(perFrame * n));
    frames = n;
    return base != nullptr;
  }
  void release() {
    std::free(base);
    base = nullptr;
  }
  uint16_t* label(uint32_t f) const { return reinterpret_cast<uint16_t*>(base + perFrame * f); }
  uint8_t* bin(uint32_t f) const { return base + perFrame * f + labelBytes; }
  uint8_t* occluded(uint32_t f) const { return bin(f) + packedBytes; }
  uint8_t* invalid(uint32_t f) const { return bin(f) + 2 * packedBytes; }
};

// Persistent worker threads for "parallel for" over file reads.
class WorkerPool {
 public:
  explicit WorkerPool(int n) {
    for (int i = 0; i < std::max(1, n); ++i) threads_.emplace_back([this]() { work(); });
  }
  ~WorkerPool() {
    {
      std::lock_guard<std::mutex> lock(m_);
      stop_ = true;
    }
    wake_.notify_all();
    for (std::thread& t : threads_) t.join();
  }
  // fn(i) for i in [0, count); returns when all are done; the first exception is rethrown here
  void run(size_t count, const std::function<void(size_t)>& fn) {
    if (count == 0) return;
    std::unique_lock<std::mutex> lock(m_);
    fn_ = &fn;
    count_ = count;
    next_ = 0;
    pending_ = count;
    error_ = nullptr;
    ++generation_;
    wake_.notify_all();
    done_.wait(lock, [this]() { return pending_ == 0; });
    fn_ = nullptr;
    if (error_) std::rethrow_exception(error_);
  }

 private:
  void work() {
    uint64_t seen = 0;
    std::unique_lock<std::mutex> lock(m_);
    for (;;) {
      wake_.wait(lock, [&]() { return stop_ || (generation_ != seen && next_ < count_); });
      if (stop_) return;
      while (next_ < count_) {
        const size_t i = next_++;
        const std::function<void(size_t)>* fn = fn_;
        lock.unlock();
        std::exception_ptr err;
        try {
          (*fn)(i);
        } catch (...) {
          err = std::current_exception();
        }
        lock.lock();
        if (err && !error_) error_ = err;
        if (--pending_ == 0) done_.notify_all();
      }
      seen = generation_;
    }
  }
  std::vector<std::thread> threads_;
  std::mutex m_;
  std::condition_variable wake_, done_;
  const std::function<void(size_t)>* fn_ = nullptr;
  size_t count_ = 0, next_ = 0, pending_ = 0;
  uint64_t generation_ = 0;
  bool stop_ = false;
  std::exception_ptr error_;
};

// Disk -> pinned staging ring -> device.  The ring has two halves of kChunk slots: the workers read the files of
// one chunk into one half while the uploads of the other half are in flight (vx_upload_ticket / vx_upload_wait).
class ScanLoader {
 public:
  static constexpr uint32_t kChunk = 16;
  ScanLoader(const KittiReader& reader, vx_ctx* ctx, int threads) : reader_(reader), ctx_(ctx), pool_(threads) {}
  ~ScanLoader() {
    if (ring_) vx_host_free(ring_);
  }
  // pinned staging ring with slots for the largest scan of the sequence (sizes from the directory, no file is opened)
  bool prepare() {
    uint64_t largest = 0;
    for (uint32_t t = 0; t < reader_.count(); ++t) {
      struct stat st;
      if (::stat(reader_.velodyneFilename(t).c_str(), &st) == 0) largest = std::max<uint64_t>(largest, uint64_t(st.st_size));
    }
    capacity_ = uint32_t(largest / 16) + 1;
    slotBytes_ = ((size_t(capacity_) * 20 + 255) / 256) * 256;
    ring_ = static_cast<uint8_t*>(vx_host_alloc(slotBytes_ * 2 * kChunk));
    return ring_ != nullptr;
  }
  // uploads the scans `todo` and records them in `resident`; throws std::runtime_error
  void load(const std::vector<uint32_t>& todo, std::map<uint32_t, ResidentScan>& resident) {
    if (todo.empty()) return;
    uint64_t ticket[2] = {0, 0};
    uint32_t points[kChunk], labelled[kChunk];
    bool fits[kChunk];
    for (size_t base = 0, turn = 0; base < todo.size(); base += kChunk, ++turn) {
      const size_t n = std::min<size_t>(kChunk, todo.size() - base);
      const size_t half = turn & 1;
      if (turn >= 2 && vx_upload_wait(ctx_, ticket[half]) != VX_OK) throw std::runtime_error(vx_last_error(ctx_));
      pool_.run(n, [&](size_t i) { fits[i] = reader_.loadInto(todo[base + i], xyzr(half, i), labels(half, i), capacity_, points[i], labelled[i]); });
      for (size_t i = 0; i < n; ++i) {
        const uint32_t t = todo[base + i];
        ResidentScan r;
        r.pose = reader_.pose(t);
        if (fits[i]) {
          r.points = points[i];
          r.labelled = labelled[i];
          if (vx_upload_scan(ctx_, t, xyzr(half, i), labels(half, i), r.points) != VX_OK) throw std::runtime_error(vx_last_error(ctx_));
        } else {  // a file that grew since the ring was sized: the plain path
          Laserscan scan;
          std::vector<uint32_t> lab;
          reader_.load(t, scan, lab);
          r.points = scan.size();
          for (const uint32_t l : lab) r.labelled += l > 0 ? 1u : 0u;
          if (vx_upload_scan(ctx_, t, scan.xyzr.data(), lab.data(), r.points) != VX_OK) throw std::runtime_error(vx_last_error(ctx_));
        }
        resident[t] = r;
      }
      ticket[half] = vx_upload_ticket(ctx_);
    }
    // the ring may be reused (or freed) by the next call only after these uploads have landed
    if (vx_upload_wait(ctx_, vx_upload_ticket(ctx_)) != VX_OK) throw std::runtime_error(vx_last_error(ctx_));
  }

 private:
  float* xyzr(size_t half, size_t i) const { return reinterpret_cast<float*>(ring_ + slotBytes_ * (half * kChunk + i)); }
  uint32_t* labels(size_t half, size_t i) const { return reinterpret_cast<uint32_t*>(ring_ + slotBytes_ * (half * kChunk + i) + size_t(capacity_) * 16); }
  const KittiReader& reader_;
  vx_ctx* ctx_;
  WorkerPool pool_;
  uint8_t* ring_ = nullptr;
  size_t slotBytes_ = 0;
  uint32_t capacity_ = 0;
};

}  // namespace

std::vector<uint32_t> planTargets(uint32_t scanCount, const std::vector<Matrix4f>& poses, const Config& config) {
  std::vector<uint32_t> targets;
  uint32_t current = 0;
  while (current < scanCount) {
    targets.push_back(current);
    float distance = 0.0f;
    uint32_t count = 0;
    auto wantMore = [&]() { return count < config.stride_num || distance < config.stride_distance || count == 0; };
    while (wantMore() && size_t(current) + 1 < poses.size()) {
      distance += poseDistance(poses[current], poses[current + 1]);
      ++count;
      ++current;
    }
    if (wantMore() && size_t(current) + 1 >= poses.size()) break;  // no further scan can be extracted
  }
  return targets;
}

void frameTransforms(const Matrix4f& anchor, const Matrix4f* poses, uint32_t n, float* ap16, float* endpoints) {
  const Matrix4f anchorInv = anchor.inverse();
  for (uint32_t s = 0; s < n; ++s) {
    const Matrix4f ap = anchorInv * poses[s];
    std::memcpy(ap16 + 16 * size_t(s), ap.m, sizeof(ap.m));
    if (endpoints) {
      endpoints[3 * size_t(s) + 0] = ap.m[12];
      endpoints[3 * size_t(s) + 1] = ap.m[13];
      endpoints[3 * size_t(s) + 2] = ap.m[14];
    }
  }
}

void describe(const Config& config, vx_grid_desc& grid, vx_filter_desc& filter, std::vector<uint32_t>& scratch) {
  grid.resolution = config.voxelSize;
  for (int a = 0; a < 3; ++a) {
    grid.min_extent[a] = config.minExtent[a];
    grid.max_extent[a] = config.maxExtent[a];
  }
  scratch.clear();
  for (const auto& join : config.joinedLabels)
    for (const uint32_t member : join.second) {
      scratch.push_back(join.first);
      scratch.push_back(member);
    }
  filter.min_range = config.minRange;
  filter.max_range = config.maxRange;
  filter.hidecar = config.hidecar ? 1 : 0;
  filter.filtered_labels = config.filteredLabels.empty() ? nullptr : config.filteredLabels.data();
  filter.n_filtered = static_cast<uint32_t>(config.filteredLabels.size());
  filter.join_pairs = scratch.empty() ? nullptr : scratch.data();
  filter.n_join_pairs = static_cast<uint32_t>(scratch.size() / 2);
}

int runGenData(const Config& config, const std::string& sequenceDir, const std::string& outputDir,
               const GenDataOptions& opt, GenDataStats* stats, std::string* error) {
  auto failWith = [&](const std::string& msg) {
    if (error) *error = msg;
    return -1;
  };
  const double t_begin = now();
  GenDataStats st;

  if (opt.writeFiles && !makePath(outputDir)) return failWith("Unable to create output directory.");

  KittiReader reader;
  reader.setVerbose(false);
  try {
    reader.initialize(sequenceDir);
  } catch (const std::exception& e) {
    return failWith(e.what());
  }
  reader.setNumPriorScans(int32_t(config.priorScans));
  reader.setNumPastScans(int32_t(config.pastScans));
  const std::vector<Matrix4f> poses = reader.poses();

  // ---- plan: targets of this shard
  std::vector<uint32_t> targets = planTargets(reader.count(), poses, config);
  {
    const size_t total = targets.size();
    const size_t lo = total * opt.shardIndex / std::max<uint32_t>(1, opt.shardCount);
    const size_t hi = total * (size_t(opt.shardIndex) + 1) / std::max<uint32_t>(1, opt.shardCount);
    std::vector<uint32_t> mine(targets.begin() + long(lo), targets.begin() + long(hi));
    if (opt.firstFrame > 0) mine.erase(mine.begin(), mine.begin() + long(std::min<size_t>(mine.size(), size_t(opt.firstFrame))));
    if (opt.maxFrames >= 0 && mine.size() > size_t(opt.maxFrames)) mine.resize(size_t(opt.maxFrames));
    targets.swap(mine);
  }

  // ---- backend
  vx_grid_desc gd;
  vx_filter_desc fd;
  std::vector<uint32_t> joinScratch;
  describe(config, gd, fd, joinScratch);
  vx_options vo;
  std::memset(&vo, 0, sizeof(vo));
  vo.frames_in_flight = opt.framesInFlight;
  vx_ctx* ctx = nullptr;
  if (vx_create(opt.device, &gd, &fd, &vo, &ctx) != VX_OK) return failWith(std::string("vx_create: ") + vx_last_global_error());
  vx_grid_info gi;
  vx_grid_info_get(ctx, &gi);
  if (opt.verbose)
    std::cout << "[voxelizer_b200] " << gi.resolution << "; num. voxels = [" << gi.size[0] << ", " << gi.size[1] << ", " << gi.size[2]
              << "], " << targets.size() << " target scans, batches of " << vx_frames_in_flight(ctx) << std::endl;

  // equal batches: a short last batch would leave most of the device idle for a whole batch time
  uint32_t batch = std::max<uint32_t>(1, vx_frames_in_flight(ctx));
  if (!targets.empty()) {
    const size_t n_batches = (targets.size() + batch - 1) / batch;
    batch = uint32_t((targets.size() + n_batches - 1) / n_batches);
  }
  OutputRing rings[2];
  std::vector<std::thread> writers;
  std::atomic<bool> writeFailed(false);
  auto joinWriters = [&]() {
    for (std::thread& t : writers) t.join();
    writers.clear();
  };
  std::map<uint32_t, ResidentScan> resident;
  ScanLoader loader(reader, ctx, std::max(1, opt.readThreads));
  int rc = 0;
  std::string err;
  if (!targets.empty() && !loader.prepare()) {
    rc = -1;
    err = "pinned staging ring allocation failed";
  }

  for (size_t base = 0, turn = 0; base < targets.size() && rc == 0; base += batch, ++turn) {
    const uint32_t nb = uint32_t(std::min<size_t>(batch, targets.size() - base));
    OutputRing& ring = rings[turn & 1];

    // ---- residency: scans [lo, hi) of this batch; older ones are evicted
    int32_t lo = INT32_MAX, hi = 0;
    for (uint32_t f = 0; f < nb; ++f) {
      int32_t pb, pe, qb, qe;
      reader.window(int32_t(targets[base + f]), pb, pe, qb, qe);
      lo = std::min(lo, std::min(pb, qb));
      hi = std::max(hi, std::max(pe, qe));
    }
    const double t_read = now();
    for (auto it = resident.begin(); it != resident.end();) {
      if (int32_t(it->first) < lo) {
        vx_evict_scan(ctx, it->first);
        it = resident.erase(it);
      } else {
        ++it;
      }
    }
    std::vector<uint32_t> todo;
    for (int32_t t = lo; t < hi; ++t)
      if (!resident.count(uint32_t(t))) todo.push_back(uint32_t(t));
    try {
      loader.load(todo, resident);
    } catch (const std::exception& e) {
      rc = -1;
      err = e.what();
    }
    if (rc == 0 && opt.writeFiles && !ring.base && !ring.allocate(std::min<uint32_t>(batch, uint32_t(targets.size())), gi)) {
      rc = -1;
      err = "output ring allocation failed";
    }
    if (rc != 0) break;
    st.read_s += now() - t_read;

    // ---- frame descriptors
    std::vector<vx_frame_desc> descs;
    std::vector<uint32_t> descTarget;
    std::vector<std::vector<uint32_t>> ids(size_t(nb) * 2);
    std::vector<std::vector<float>> mats(size_t(nb) * 2), eps(nb);
    for (uint32_t f = 0; f < nb; ++f) {
      const uint32_t target = targets[base + f];
      int32_t pb, pe, qb, qe;
      reader.window(int32_t(target), pb, pe, qb, qe);
      uint32_t labelCount = 0, pointCount = 0;  // gen_data.cpp:75-86
      for (int32_t t = qb; t < qe; ++t) {
        pointCount += resident[uint32_t(t)].points;
        labelCount += resident[uint32_t(t)].labelled;
      }
      const float percentageLabeled = 100.0f * float(labelCount) / float(pointCount);
      if (!(percentageLabeled > 0.0f)) {
        if (opt.verbose) std::cout << "skipped." << std::endl;
        ++st.skipped;
        continue;
      }
      const Matrix4f anchor = resident[uint32_t(pe - 1)].pose;  // pose of the last prior scan = the target
      std::vector<Matrix4f> pp, qp;
      for (int32_t t = pb; t < pe; ++t) {
        ids[2 * f].push_back(uint32_t(t));
        pp.push_back(resident[uint32_t(t)].pose);
      }
      for (int32_t t = qb; t < qe; ++t) {
        ids[2 * f + 1].push_back(uint32_t(t));
        qp.push_back(resident[uint32_t(t)].pose);
      }
      mats[2 * f].resize(pp.size() * 16);
      mats[2 * f + 1].resize(qp.size() * 16);
      eps[f].resize(qp.size() * 3);
      frameTransforms(anchor, pp.data(), uint32_t(pp.size()), mats[2 * f].data(), nullptr);
      frameTransforms(anchor, qp.data(), uint32_t(qp.size()), mats[2 * f + 1].data(), eps[f].data());
      vx_frame_desc d;
      std::memset(&d, 0, sizeof(d));
      d.n_prior = uint32_t(pp.size());
      d.prior_ids = ids[2 * f].data();
      d.ap_prior = mats[2 * f].data();
      d.n_past = uint32_t(qp.size());
      d.past_ids = ids[2 * f + 1].data();
      d.ap_past = mats[2 * f + 1].data();
      d.n_endpoints = uint32_t(qp.size());
      d.endpoints = eps[f].data();
      if (opt.writeFiles) {
        const uint32_t k = uint32_t(descs.size());
        d.out_label = ring.label(k);
        d.out_bin = ring.bin(k);
        d.out_occluded = ring.occluded(k);
        d.out_invalid = ring.invalid(k);
      }
      descs.push_back(d);
      descTarget.push_back(target);
    }

    // the ring we are about to overwrite was handed to the writers two batches ago
    const double t_wait = now();
    joinWriters();
    st.write_s += now() - t_wait;
    if (writeFailed) {
      rc = -1;
      err = "failed to write an output file";
      break;
    }

    const double t_gpu = now();
    if (!descs.empty() && vx_process_frames(ctx, descs.data(), uint32_t(descs.size())) != VX_OK) {
      rc = -1;
      err = std::string("vx_process_frames: ") + vx_last_error(ctx);
      break;
    }
    st.gpu_s += now() - t_gpu;
    st.frames += descs.size();

    if (opt.writeFiles) {
      const int nthreads = std::max(1, opt.ioThreads);
      const size_t nf = descs.size();
      for (int w = 0; w < nthreads; ++w) {
        writers.emplace_back([&, w, nthreads, nf, descTarget, ring]() {
          for (size_t k = size_t(w); k < nf; k += size_t(nthreads)) {
            char name[32];
            std::snprintf(name, sizeof(name), "%06u", descTarget[k]);  // setfill('0') << setw(6)
            const std::string stem = outputDir + "/" + name;
            bool ok = writeAll(stem + ".bin", ring.bin(uint32_t(k)), ring.packedBytes);
            ok = writeAll(stem + ".label", ring.label(uint32_t(k)), ring.labelBytes) && ok;
            ok = writeAll(stem + ".occluded", ring.occluded(uint32_t(k)), ring.packedBytes) && ok;
            ok = writeAll(stem + ".invalid", ring.invalid(uint32_t(k)), ring.packedBytes) && ok;
            if (!ok) writeFailed = true;
          }
        });
      }
    }
  }
  const double t_tail = now();
  joinWriters();
  st.write_s += now() - t_tail;
  if (rc == 0 && writeFailed) {
    rc = -1;
    err = "failed to write an output file";
  }
  vx_stage_times_get(ctx, &st.stages);
  rings[0].release();
  rings[1].release();
  vx_destroy(ctx);
  st.wall_s = now() - t_begin;
  if (stats) *stats = st;
  if (rc != 0) return failWith(err);
  return 0;
}

}  // namespace vxb
