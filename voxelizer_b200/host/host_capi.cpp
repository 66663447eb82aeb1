// C entry points of the host layer for non-C++ callers (the Python binding in voxelizer_b200/__init__.py,
// tests, bench.py).  Thin marshalling over config.h / kitti_reader.h / gen_data_pipeline.h.
#include <cstring>
#include <string>

#include "gen_data_pipeline.h"
#include "voxel_grid.h"

namespace {
std::string g_error;
}

extern "C" {

const char* vxh_last_error() { return g_error.c_str(); }

struct vxh_config_pod {
  float voxelSize;
  float minExtent[4];
  float maxExtent[4];
  uint32_t maxNumScans, priorScans, pastScans;
  float pastDistance, minRange, maxRange;
  uint32_t stride_num;
  float stride_distance;
  int32_t hidecar;
  uint32_t n_filtered;
  uint32_t n_join_pairs;
};

void* vxh_config_load(const char* path) {
  try {
    return new vxb::Config(vxb::parseConfiguration(path));
  } catch (const std::exception& e) {
    g_error = e.what();
    return nullptr;
  }
}

void vxh_config_free(void* c) { delete static_cast<vxb::Config*>(c); }

void vxh_config_get(void* c, vxh_config_pod* o) {
  const vxb::Config& cfg = *static_cast<vxb::Config*>(c);
  o->voxelSize = cfg.voxelSize;
  for (int i = 0; i < 4; ++i) {
    o->minExtent[i] = cfg.minExtent[i];
    o->maxExtent[i] = cfg.maxExtent[i];
  }
  o->maxNumScans = cfg.maxNumScans;
  o->priorScans = cfg.priorScans;
  o->pastScans = cfg.pastScans;
  o->pastDistance = cfg.pastDistance;
  o->minRange = cfg.minRange;
  o->maxRange = cfg.maxRange;
  o->stride_num = cfg.stride_num;
  o->stride_distance = cfg.stride_distance;
  o->hidecar = cfg.hidecar ? 1 : 0;
  o->n_filtered = uint32_t(cfg.filteredLabels.size());
  uint32_t n = 0;
  for (const auto& j : cfg.joinedLabels) n += uint32_t(j.second.size());
  o->n_join_pairs = n;
}

void vxh_config_filtered(void* c, uint32_t* out) {
  const vxb::Config& cfg = *static_cast<vxb::Config*>(c);
  for (size_t i = 0; i < cfg.filteredLabels.size(); ++i) out[i] = cfg.filteredLabels[i];
}

void vxh_config_joins(void* c, uint32_t* out_pairs) {
  const vxb::Config& cfg = *static_cast<vxb::Config*>(c);
  size_t n = 0;
  for (const auto& j : cfg.joinedLabels)
    for (const uint32_t member : j.second) {
      out_pairs[2 * n] = j.first;
      out_pairs[2 * n + 1] = member;
      ++n;
    }
}

void vxh_mat_inverse(const float* a16, float* out16) {
  vxb::Matrix4f a;
  std::memcpy(a.m, a16, 64);
  const vxb::Matrix4f r = a.inverse();
  std::memcpy(out16, r.m, 64);
}

void vxh_mat_mul(const float* a16, const float* b16, float* out16) {
  vxb::Matrix4f a, b;
  std::memcpy(a.m, a16, 64);
  std::memcpy(b.m, b16, 64);
  const vxb::Matrix4f r = a * b;
  std::memcpy(out16, r.m, 64);
}

// ap16[n x 16], endpoints[n x 3] (nullable) for poses16[n x 16] relative to anchor16
void vxh_frame_transforms(const float* anchor16, const float* poses16, uint32_t n, float* ap16, float* endpoints) {
  vxb::Matrix4f anchor;
  std::memcpy(anchor.m, anchor16, 64);
  vxb::frameTransforms(anchor, reinterpret_cast<const vxb::Matrix4f*>(poses16), n, ap16, endpoints);
}

void* vxh_reader_open(const char* dir) {
  try {
    vxb::KittiReader* r = new vxb::KittiReader;
    r->setVerbose(false);
    r->initialize(dir);
    return r;
  } catch (const std::exception& e) {
    g_error = e.what();
    return nullptr;
  }
}
void vxh_reader_free(void* r) { delete static_cast<vxb::KittiReader*>(r); }
uint32_t vxh_reader_count(void* r) { return static_cast<vxb::KittiReader*>(r)->count(); }
uint32_t vxh_reader_num_poses(void* r) { return uint32_t(static_cast<vxb::KittiReader*>(r)->poses().size()); }
void vxh_reader_poses(void* r, float* out16n) {
  const std::vector<vxb::Matrix4f> p = static_cast<vxb::KittiReader*>(r)->poses();
  for (size_t i = 0; i < p.size(); ++i) std::memcpy(out16n + 16 * i, p[i].m, 64);
}
// window of target idx for the given prior/past settings: out4 = priorBegin, priorEnd, pastBegin, pastEnd
void vxh_reader_window(void* r, int32_t prior, int32_t past, int32_t idx, int32_t* out4) {
  vxb::KittiReader* k = static_cast<vxb::KittiReader*>(r);
  k->setNumPriorScans(prior);
  k->setNumPastScans(past);
  k->window(idx, out4[0], out4[1], out4[2], out4[3]);
}

// One scan through either loader: direct = KittiReader::loadInto (raw labels, straight into caller memory), else
// KittiReader::load (labels masked with 0xFFFF).  Returns 0, 1 = does not fit `capacity`, -1 = error (vxh_last_error).
int vxh_reader_load(void* r, uint32_t t, int direct, float* xyzr, uint32_t* labels, uint32_t capacity, uint32_t* points, uint32_t* labelled) {
  try {
    const vxb::KittiReader* k = static_cast<vxb::KittiReader*>(r);
    if (direct) return k->loadInto(t, xyzr, labels, capacity, *points, *labelled) ? 0 : 1;
    vxb::Laserscan scan;
    std::vector<uint32_t> lab;
    k->load(t, scan, lab);
    if (scan.size() > capacity) return 1;
    *points = scan.size();
    *labelled = 0;
    for (const uint32_t l : lab) *labelled += l > 0 ? 1u : 0u;
    if (*points) {
      std::memcpy(xyzr, scan.xyzr.data(), size_t(*points) * 16);
      std::memcpy(labels, lab.data(), size_t(*points) * 4);
    }
    return 0;
  } catch (const std::exception& e) {
    g_error = e.what();
    return -1;
  }
}

// target list (gen_data.cpp:62,129-141).  Returns the number of targets; writes at most cap.
uint32_t vxh_plan_targets(void* cfg, uint32_t scan_count, const float* poses16, uint32_t n_poses, uint32_t* out, uint32_t cap) {
  std::vector<vxb::Matrix4f> poses(n_poses);
  if (n_poses) std::memcpy(poses.data(), poses16, size_t(n_poses) * 64);
  const std::vector<uint32_t> t = vxb::planTargets(scan_count, poses, *static_cast<vxb::Config*>(cfg));
  for (size_t i = 0; i < t.size() && i < cap; ++i) out[i] = t[i];
  return uint32_t(t.size());
}

struct vxh_gen_data_stats {
  uint64_t frames, skipped;
  double wall_s, read_s, gpu_s, write_s;
  vx_stage_times stages;
};

int vxh_gen_data(const char* cfg_path, const char* seq_dir, const char* out_dir, int device, uint32_t shard_index,
                 uint32_t shard_count, uint32_t frames_in_flight, int64_t first_frame, int64_t max_frames, int write_files,
                 int verbose, vxh_gen_data_stats* stats) {
  try {
    const vxb::Config cfg = vxb::parseConfiguration(cfg_path);
    vxb::GenDataOptions opt;
    opt.device = device;
    opt.shardIndex = shard_index;
    opt.shardCount = shard_count ? shard_count : 1;
    opt.framesInFlight = frames_in_flight;
    opt.firstFrame = first_frame;
    opt.maxFrames = max_frames;
    opt.writeFiles = write_files != 0;
    opt.verbose = verbose != 0;
    vxb::GenDataStats st;
    std::string err;
    const int rc = vxb::runGenData(cfg, seq_dir, out_dir, opt, &st, &err);
    if (stats) {
      stats->frames = st.frames;
      stats->skipped = st.skipped;
      stats->wall_s = st.wall_s;
      stats->read_s = st.read_s;
      stats->gpu_s = st.gpu_s;
      stats->write_s = st.write_s;
      stats->stages = st.stages;
    }
    if (rc != 0) g_error = err;
    return rc;
  } catch (const std::exception& e) {
    g_error = e.what();
    return -1;
  }
}

}  // extern "C"
