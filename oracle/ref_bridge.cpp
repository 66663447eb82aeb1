// ORACLE / TEST INFRASTRUCTURE ONLY -- the product never links, loads or calls this.
//
// C-ABI bridge over the REFERENCE's own code.  Built by `make -C oracle ref` together with the
// reference translation units compiled unmodified from /root/reference/src (VoxelGrid.cpp,
// voxelize_utils.cpp, kitti_utils.cpp, KittiReader.cpp, rv/string_utils.cpp) into
// oracle/_ref/libvx_ref.so.  Third-party headers the reference needs and this image lacks are the
// stand-ins under oracle/shim/ (see their headers for what they restate).  This file contains no
// algorithm: it only marshals flat arrays into the reference's types and calls
//   parseConfiguration / fillVoxelGrid / saveVoxelGrid   (src/data/voxelize_utils.h:31-48)
//   VoxelGrid::{initialize,clear,insert,updateOcclusions,updateInvalid,isOccluded,isInvalid}
//                                                         (src/data/VoxelGrid.h:26-77)
//   KittiReader::{initialize,count,poses}                 (src/widget/KittiReader.h:21-35)
// Matrices cross this ABI as 16 floats, column-major (Eigen::Matrix4f storage order).
#include <cstdint>
#include <cstring>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>

#include "data/voxelize_utils.h"
#include "widget/KittiReader.h"

namespace {

class RefGrid : public VoxelGrid {
 public:
  const std::vector<int32_t>& occlusions() const { return occlusions_; }
  const std::vector<int32_t>& invalid() const { return invalid_; }
  const std::vector<int32_t>& memo() const { return occludedBy_; }
};

Eigen::Matrix4f to_mat(const float* m16) {
  Eigen::Matrix4f m;
  std::memcpy(m.data(), m16, 16 * sizeof(float));
  return m;
}

// The reference prints banners to std::cout; keep test output clean.
struct Quiet {
  std::streambuf* old;
  std::ostringstream sink;
  Quiet() : old(std::cout.rdbuf(sink.rdbuf())) {}
  ~Quiet() { std::cout.rdbuf(old); }
};

std::string g_error;

}  // namespace

extern "C" {

const char* vxref_last_error() { return g_error.c_str(); }

// ---------------------------------------------------------------- Config
struct vxref_config_pod {
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

void* vxref_config_load(const char* path) {
  try {
    Quiet q;
    return new Config(parseConfiguration(path));
  } catch (const std::exception& e) {
    g_error = e.what();
    return nullptr;
  }
}

void vxref_config_free(void* c) { delete static_cast<Config*>(c); }

void vxref_config_get(void* c, vxref_config_pod* out) {
  const Config& cfg = *static_cast<Config*>(c);
  out->voxelSize = cfg.voxelSize;
  for (int i = 0; i < 4; ++i) {
    out->minExtent[i] = cfg.minExtent[i];
    out->maxExtent[i] = cfg.maxExtent[i];
  }
  out->maxNumScans = cfg.maxNumScans;
  out->priorScans = cfg.priorScans;
  out->pastScans = cfg.pastScans;
  out->pastDistance = cfg.pastDistance;
  out->minRange = cfg.minRange;
  out->maxRange = cfg.maxRange;
  out->stride_num = cfg.stride_num;
  out->stride_distance = cfg.stride_distance;
  out->hidecar = cfg.hidecar ? 1 : 0;
  out->n_filtered = uint32_t(cfg.filteredLabels.size());
  uint32_t n = 0;
  for (const auto& j : cfg.joinedLabels) n += uint32_t(j.second.size());
  out->n_join_pairs = n;
}

void vxref_config_filtered(void* c, uint32_t* out) {
  const Config& cfg = *static_cast<Config*>(c);
  for (size_t i = 0; i < cfg.filteredLabels.size(); ++i) out[i] = cfg.filteredLabels[i];
}

// pairs (key, member) in map order
void vxref_config_joins(void* c, uint32_t* out_pairs) {
  const Config& cfg = *static_cast<Config*>(c);
  size_t n = 0;
  for (const auto& j : cfg.joinedLabels)
    for (uint32_t v : j.second) {
      out_pairs[2 * n] = j.first;
      out_pairs[2 * n + 1] = v;
      ++n;
    }
}

// ---------------------------------------------------------------- VoxelGrid
void* vxref_grid_new(float resolution, const float* min4, const float* max4) {
  Quiet q;
  RefGrid* g = new RefGrid;
  g->initialize(resolution, Eigen::Vector4f(min4[0], min4[1], min4[2], min4[3]),
                Eigen::Vector4f(max4[0], max4[1], max4[2], max4[3]));
  return g;
}

void vxref_grid_free(void* g) { delete static_cast<RefGrid*>(g); }

void vxref_grid_dims(void* gp, uint32_t* size3, float* offset4, float* resolution) {
  RefGrid& g = *static_cast<RefGrid*>(gp);
  for (uint32_t d = 0; d < 3; ++d) size3[d] = g.size(d);
  for (int i = 0; i < 4; ++i) offset4[i] = g.offset()[i];
  *resolution = g.resolution();
}

void vxref_grid_clear(void* gp) { static_cast<RefGrid*>(gp)->clear(); }

void vxref_grid_insert(void* gp, const float* p4, const uint32_t* labels, uint64_t n) {
  RefGrid& g = *static_cast<RefGrid*>(gp);
  for (uint64_t i = 0; i < n; ++i)
    g.insert(Eigen::Vector4f(p4[4 * i], p4[4 * i + 1], p4[4 * i + 2], p4[4 * i + 3]), labels[i]);
}

// fillVoxelGrid(anchor, scans, labels, grid, cfg).  xyzr: per scan n x 4 floats (KITTI .bin layout);
// labels are used as given (the reader's & 0xFFFF is the caller's job, as in the reference).
void vxref_fill(void* gp, void* cfgp, const float* anchor16, uint32_t nscans, const float* const* xyzr,
                const uint32_t* const* labels, const uint32_t* counts, const float* poses16) {
  RefGrid& g = *static_cast<RefGrid*>(gp);
  const Config& cfg = *static_cast<Config*>(cfgp);
  std::vector<PointcloudPtr> pts;
  std::vector<LabelsPtr> labs;
  for (uint32_t s = 0; s < nscans; ++s) {
    PointcloudPtr pc(new Laserscan);
    pc->pose = to_mat(poses16 + 16 * size_t(s));
    pc->points.resize(counts[s]);
    pc->remissions.resize(counts[s]);
    for (uint32_t i = 0; i < counts[s]; ++i) {
      pc->points[i].x = xyzr[s][4 * size_t(i) + 0];
      pc->points[i].y = xyzr[s][4 * size_t(i) + 1];
      pc->points[i].z = xyzr[s][4 * size_t(i) + 2];
      pc->remissions[i] = xyzr[s][4 * size_t(i) + 3];
    }
    pts.push_back(pc);
    labs.push_back(LabelsPtr(new std::vector<uint32_t>(labels[s], labels[s] + counts[s])));
  }
  fillVoxelGrid(to_mat(anchor16), pts, labs, g, cfg);
}

void vxref_update_occlusions(void* gp) { static_cast<RefGrid*>(gp)->updateOcclusions(); }
void vxref_insert_occlusion_labels(void* gp) { static_cast<RefGrid*>(gp)->insertOcclusionLabels(); }
// VoxelGrid::occludedBy as the public method (VoxelGrid.h:89-90)
int32_t vxref_occluded_by(void* gp, int32_t i, int32_t j, int32_t k, const float* e3, int32_t* visited, uint32_t capacity, uint32_t* n_visited) {
  std::vector<Eigen::Vector3i> seen;
  const int32_t r = static_cast<RefGrid*>(gp)->occludedBy(i, j, k, Eigen::Vector3f(e3[0], e3[1], e3[2]), &seen);
  if (n_visited) *n_visited = uint32_t(seen.size());
  for (uint32_t q = 0; q < seen.size() && q < capacity; ++q)
    for (int a = 0; a < 3; ++a) visited[3 * q + uint32_t(a)] = seen[q][a];
  return r;
}

void vxref_update_invalid(void* gp, const float* e3) {
  static_cast<RefGrid*>(gp)->updateInvalid(Eigen::Vector3f(e3[0], e3[1], e3[2]));
}

// what: 0 = occlusions_, 1 = invalid_, 2 = occludedBy_; reference-internal index i + j*Sx + k*Sx*Sy.
uint64_t vxref_grid_dump(void* gp, int what, int32_t* out) {
  RefGrid& g = *static_cast<RefGrid*>(gp);
  const std::vector<int32_t>& v = what == 0 ? g.occlusions() : (what == 1 ? g.invalid() : g.memo());
  if (out) std::memcpy(out, v.data(), v.size() * sizeof(int32_t));
  return v.size();
}

// per-voxel point count, reference-internal index order
void vxref_grid_counts(void* gp, uint32_t* out) {
  RefGrid& g = *static_cast<RefGrid*>(gp);
  const auto& vox = g.voxels();
  for (size_t i = 0; i < vox.size(); ++i) out[i] = vox[i].count;
}

// histogram as (internal voxel index, label, count) triples in (voxel, label) order; returns the
// number of triples (call with out == nullptr to size).
uint64_t vxref_grid_hist(void* gp, uint32_t* out) {
  RefGrid& g = *static_cast<RefGrid*>(gp);
  const auto& vox = g.voxels();
  uint64_t n = 0;
  for (size_t i = 0; i < vox.size(); ++i)
    for (const auto& kv : vox[i].labels) {
      if (out) {
        out[3 * n] = uint32_t(i);
        out[3 * n + 1] = kv.first;
        out[3 * n + 2] = kv.second;
      }
      ++n;
    }
  return n;
}

// isOccluded / isInvalid in the saveVoxelGrid output order (x, then y, then z fastest).
void vxref_grid_flags(void* gp, uint8_t* occluded, uint8_t* invalid) {
  RefGrid& g = *static_cast<RefGrid*>(gp);
  size_t n = 0;
  for (uint32_t x = 0; x < g.size(0); ++x)
    for (uint32_t y = 0; y < g.size(1); ++y)
      for (uint32_t z = 0; z < g.size(2); ++z, ++n) {
        occluded[n] = g.isOccluded(x, y, z) ? 1 : 0;
        invalid[n] = g.isInvalid(x, y, z) ? 1 : 0;
      }
}

void vxref_grid_save(void* gp, const char* dir, const char* basename, const char* mode) {
  saveVoxelGrid(*static_cast<RefGrid*>(gp), dir, basename, mode);
}

// ---------------------------------------------------------------- pose algebra (shimmed Eigen)
void vxref_mat_inverse(const float* a16, float* out16) {
  const Eigen::Matrix4f r = to_mat(a16).inverse();
  std::memcpy(out16, r.data(), 16 * sizeof(float));
}

void vxref_mat_mul(const float* a16, const float* b16, float* out16) {
  const Eigen::Matrix4f r = to_mat(a16) * to_mat(b16);
  std::memcpy(out16, r.data(), 16 * sizeof(float));
}

// gen_data.cpp:114: (anchor_pose.inverse() * pose).col(3).head(3)
void vxref_endpoint(const float* anchor16, const float* pose16, float* e3) {
  const Eigen::Vector3f e = (to_mat(anchor16).inverse() * to_mat(pose16)).col(3).head(3);
  e3[0] = e[0];
  e3[1] = e[1];
  e3[2] = e[2];
}

// ---------------------------------------------------------------- KittiReader
void* vxref_reader_open(const char* dir) {
  try {
    Quiet q;
    KittiReader* r = new KittiReader;
    r->initialize(QString::fromStdString(dir));
    return r;
  } catch (const std::exception& e) {
    g_error = e.what();
    return nullptr;
  }
}

void vxref_reader_free(void* r) { delete static_cast<KittiReader*>(r); }
uint32_t vxref_reader_count(void* r) { return static_cast<KittiReader*>(r)->count(); }
uint32_t vxref_reader_num_poses(void* r) { return uint32_t(static_cast<KittiReader*>(r)->poses().size()); }
void vxref_reader_poses(void* r, float* out16n) {
  const std::vector<Eigen::Matrix4f> p = static_cast<KittiReader*>(r)->poses();
  for (size_t i = 0; i < p.size(); ++i) std::memcpy(out16n + 16 * i, p[i].data(), 16 * sizeof(float));
}

}  // extern "C"
