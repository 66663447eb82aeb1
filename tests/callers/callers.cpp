// TEST CODE (not part of the product libraries): callers shaped like the reference's two users of the hot path,
// written against the kept host API (vxb::KittiReader, VoxelGrid, fillVoxelGrid, saveVoxelGrid, Config) to show
// that such callers run unchanged on the GPU backend:
//   * the batch program's frame loop                 (reference: src/gen_data.cpp:46-142)
//   * the GUI's rebuild of both grids + voxel lists  (reference: src/widget/Mainframe.cpp:279-332, 369-447)
//   * direct per-voxel queries and public rays       (reference: src/data/VoxelGrid.h:26-118)
// Built by tests/conftest.py into tests/callers/libvxcallers.so (links libvxhost.so).
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <stdexcept>
#include <string>

#include "gen_data_pipeline.h"
#include "voxel_grid.h"

namespace {

std::string g_error;

using vxb::Config;
using vxb::LabelsPtr;
using vxb::Matrix4f;
using vxb::PointcloudPtr;
using vxb::Vector3f;
using vxb::Vector4f;
using vxb::VoxelGrid;

struct Window {
  std::vector<PointcloudPtr> priorPoints, pastPoints;
  std::vector<LabelsPtr> priorLabels, pastLabels;
};

// share of labelled points among the past scans, in percent (fp32, 0/0 = NaN = "not labelled")
float labelledPercent(const std::vector<LabelsPtr>& labels) {
  uint32_t labelled = 0, total = 0;
  for (const LabelsPtr& l : labels) {
    total += uint32_t(l->size());
    for (const uint32_t v : *l) labelled += v > 0 ? 1u : 0u;
  }
  return 100.0f * float(labelled) / float(total);
}

float translationDistance(const Matrix4f& a, const Matrix4f& b) {
  const float x = a(0, 3) - b(0, 3), y = a(1, 3) - b(1, 3), z = a(2, 3) - b(2, 3);
  return std::sqrt(x * x + (y * y + z * z));
}

// both grids of one target, the way either caller builds them
void buildGrids(const Config& config, const Window& w, VoxelGrid& priorGrid, VoxelGrid& pastGrid) {
  const Matrix4f anchor = w.priorPoints.back()->pose;
  fillVoxelGrid(anchor, w.priorPoints, w.priorLabels, priorGrid, config);
  fillVoxelGrid(anchor, w.priorPoints, w.priorLabels, pastGrid, config);
  fillVoxelGrid(anchor, w.pastPoints, w.pastLabels, pastGrid, config);
  priorGrid.updateOcclusions();
  pastGrid.updateOcclusions();
  const Matrix4f anchorInverse = anchor.inverse();
  for (const PointcloudPtr& scan : w.pastPoints) {
    const Matrix4f rel = anchorInverse * scan->pose;
    pastGrid.updateInvalid(Vector3f(rel(0, 3), rel(1, 3), rel(2, 3)));
  }
}

}  // namespace

extern "C" {

const char* vxt_last_error() { return g_error.c_str(); }

// The batch program, one frame at a time through VoxelGrid.  Returns 0 and the number of frames written.
int vxt_gen_data_frame_by_frame(const char* cfg_path, const char* seq_dir, const char* out_dir, int device, int64_t max_frames,
                                uint64_t* frames_written) {
  try {
    const Config config = vxb::parseConfiguration(cfg_path);
    vxb::KittiReader reader;
    reader.setVerbose(false);
    reader.initialize(seq_dir);
    reader.setNumPriorScans(int32_t(config.priorScans));
    reader.setNumPastScans(int32_t(config.pastScans));
    VoxelGrid priorGrid(device), pastGrid(device);
    priorGrid.initialize(config.voxelSize, config.minExtent, config.maxExtent);
    pastGrid.initialize(config.voxelSize, config.minExtent, config.maxExtent);
    const std::vector<Matrix4f> poses = reader.poses();
    uint64_t written = 0;
    for (uint32_t target = 0; target < reader.count();) {
      Window w;
      reader.retrieve(int32_t(target), w.priorPoints, w.priorLabels, w.pastPoints, w.pastLabels);
      priorGrid.clear();
      pastGrid.clear();
      if (labelledPercent(w.pastLabels) > 0.0f) {
        buildGrids(config, w, priorGrid, pastGrid);
        char name[16];
        std::snprintf(name, sizeof(name), "%06u", target);
        saveVoxelGrid(priorGrid, out_dir, name, "input");
        saveVoxelGrid(pastGrid, out_dir, name, "target");
        ++written;
      }
      if (max_frames >= 0 && int64_t(written) >= max_frames) break;
      // stride walk: at least one scan, then until both the scan count and the travelled distance are reached
      float travelled = 0.0f;
      uint32_t walked = 0;
      auto unsatisfied = [&] { return walked < config.stride_num || travelled < config.stride_distance || walked == 0; };
      while (unsatisfied() && size_t(target) + 1 < poses.size()) {
        travelled += translationDistance(poses[target], poses[target + 1]);
        ++walked;
        ++target;
      }
      if (unsatisfied()) break;  // the poses ran out before the stride was complete
    }
    if (frames_written) *frames_written = written;
    return 0;
  } catch (const std::exception& e) {
    g_error = e.what();
    return -1;
  }
}

// The GUI's rebuild for one target scan: both grids, then the three voxel lists it draws.
//   labeled_*: (x, y, z, majority label) of every voxel with count > 0, x-major / z-minor order   (extractLabeledVoxels)
//   occluded / invalid: linear output index (x*Sy + y)*Sz + z of every voxel with isOccluded / isInvalid, same order
// Each list is written up to its capacity; counts[] = {prior labeled, past labeled, prior occluded, past occluded, past invalid}.
int vxt_build_voxel_grids(const char* cfg_path, const char* seq_dir, int32_t target, int device, uint32_t* labeled_prior,
                          uint32_t* labeled_past, uint32_t* occluded_prior, uint32_t* occluded_past, uint32_t* invalid_past,
                          uint64_t capacity, uint64_t* counts, double* seconds) {
  try {
    const Config config = vxb::parseConfiguration(cfg_path);
    vxb::KittiReader reader;
    reader.setVerbose(false);
    reader.initialize(seq_dir);
    reader.setNumPriorScans(int32_t(config.priorScans));
    reader.setNumPastScans(int32_t(config.pastScans));
    VoxelGrid priorGrid(device), pastGrid(device);
    priorGrid.initialize(config.voxelSize, config.minExtent, config.maxExtent);
    pastGrid.initialize(config.voxelSize, config.minExtent, config.maxExtent);
    Window w;
    reader.retrieve(target, w.priorPoints, w.priorLabels, w.pastPoints, w.pastLabels);
    const auto t0 = std::chrono::steady_clock::now();
    priorGrid.clear();
    pastGrid.clear();
    if (!w.priorPoints.empty()) buildGrids(config, w, priorGrid, pastGrid);
    auto labeled = [&](const VoxelGrid& grid, uint32_t* out) {
      uint64_t n = 0;
      for (uint32_t x = 0; x < grid.size(0); ++x)
        for (uint32_t y = 0; y < grid.size(1); ++y)
          for (uint32_t z = 0; z < grid.size(2); ++z) {
            const VoxelGrid::Voxel& v = grid(x, y, z);
            if (v.count == 0) continue;
            uint32_t bestCount = 0, bestLabel = 0;
            for (const auto& lc : v.labels)
              if (lc.second > bestCount) {
                bestCount = lc.second;
                bestLabel = lc.first;
              }
            if (n < capacity) {
              out[4 * n] = x;
              out[4 * n + 1] = y;
              out[4 * n + 2] = z;
              out[4 * n + 3] = bestLabel;
            }
            ++n;
          }
      return n;
    };
    auto flagged = [&](const VoxelGrid& grid, bool invalid, uint32_t* out) {
      uint64_t n = 0;
      for (uint32_t x = 0; x < grid.size(0); ++x)
        for (uint32_t y = 0; y < grid.size(1); ++y)
          for (uint32_t z = 0; z < grid.size(2); ++z) {
            const bool flag = invalid ? grid.isInvalid(int32_t(x), int32_t(y), int32_t(z)) : grid.isOccluded(int32_t(x), int32_t(y), int32_t(z));
            if (!flag) continue;
            if (n < capacity) out[n] = (x * grid.size(1) + y) * grid.size(2) + z;
            ++n;
          }
      return n;
    };
    counts[0] = labeled(priorGrid, labeled_prior);
    counts[1] = labeled(pastGrid, labeled_past);
    counts[2] = flagged(priorGrid, false, occluded_prior);
    counts[3] = flagged(pastGrid, false, occluded_past);
    counts[4] = flagged(pastGrid, true, invalid_past);
    if (seconds) *seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    return 0;
  } catch (const std::exception& e) {
    g_error = e.what();
    return -1;
  }
}

// VoxelGrid used directly: raw insert() of n already transformed points, updateOcclusions(), one updateInvalid() per
// endpoint, then every per-voxel query of the reference API (arrays in the reference-internal index order
// i + j*Sx + k*Sx*Sy) and n_rays public occludedBy() calls: ray r starts at ray_start[3r..] towards ray_end[3r..],
// ray_occluder[r] = its result, ray_visited receives its visited list at offset ray_offset[r] (ray_offset[n_rays] = total).
int vxt_voxel_grid_probe(int device, float resolution, const float* min3, const float* max3, const float* points4, const uint32_t* labels,
                         uint32_t n, const float* endpoints3, uint32_t n_endpoints, uint32_t* size3, float* offset3, uint8_t* occluded,
                         uint8_t* is_free, uint8_t* invalid, uint32_t* counts, uint32_t* n_labels, uint32_t* majority, uint32_t n_rays,
                         const int32_t* ray_start, const float* ray_end, int32_t* ray_occluder, int32_t* ray_visited,
                         uint64_t visited_capacity, uint64_t* ray_offset, int insert_occlusion_labels, uint16_t* label_image,
                         uint8_t* bin_image, uint8_t* occluded_image, uint8_t* invalid_image) {
  try {
    VoxelGrid grid(device);
    grid.initialize(resolution, Vector4f(min3[0], min3[1], min3[2], 1), Vector4f(max3[0], max3[1], max3[2], 1));
    for (uint32_t p = 0; p < n; ++p)
      grid.insert(Vector4f(points4[4 * p], points4[4 * p + 1], points4[4 * p + 2], points4[4 * p + 3]), labels[p]);
    grid.updateOcclusions();
    if (insert_occlusion_labels) grid.insertOcclusionLabels();  // the order of the reference's (commented-out) call sites
    for (uint32_t e = 0; e < n_endpoints; ++e) grid.updateInvalid(Vector3f(endpoints3[3 * e], endpoints3[3 * e + 1], endpoints3[3 * e + 2]));
    if (label_image) std::memcpy(label_image, grid.imageLabel().data(), grid.imageLabel().size() * 2);  // what saveVoxelGrid writes
    if (bin_image) std::memcpy(bin_image, grid.imageBin().data(), grid.imageBin().size());
    if (occluded_image) std::memcpy(occluded_image, grid.imageOccluded().data(), grid.imageOccluded().size());
    if (invalid_image) std::memcpy(invalid_image, grid.imageInvalid().data(), grid.imageInvalid().size());
    if (size3)
      for (int a = 0; a < 3; ++a) size3[a] = grid.size(uint32_t(a));
    if (offset3)
      for (int a = 0; a < 3; ++a) offset3[a] = grid.offset()[a];
    for (uint32_t k = 0; k < grid.size(2); ++k)
      for (uint32_t j = 0; j < grid.size(1); ++j)
        for (uint32_t i = 0; i < grid.size(0); ++i) {
          const size_t idx = size_t(grid.index(int32_t(i), int32_t(j), int32_t(k)));
          occluded[idx] = grid.isOccluded(int32_t(i), int32_t(j), int32_t(k));
          is_free[idx] = grid.isFree(int32_t(i), int32_t(j), int32_t(k));
          invalid[idx] = grid.isInvalid(int32_t(i), int32_t(j), int32_t(k));
          const VoxelGrid::Voxel& v = grid(i, j, k);
          counts[idx] = v.count;
          n_labels[idx] = uint32_t(v.labels.size());
          uint32_t best = 0, bestc = 0;
          for (const auto& lc : v.labels)
            if (lc.second > bestc) {
              bestc = lc.second;
              best = lc.first;
            }
          majority[idx] = best;
        }
    uint64_t at = 0;
    for (uint32_t r = 0; r < n_rays; ++r) {
      std::vector<vxb::Vector3i> seen;
      ray_occluder[r] = grid.occludedBy(ray_start[3 * r], ray_start[3 * r + 1], ray_start[3 * r + 2],
                                        Vector3f(ray_end[3 * r], ray_end[3 * r + 1], ray_end[3 * r + 2]), &seen);
      ray_offset[r] = at;
      for (const vxb::Vector3i& c : seen) {
        if (at < visited_capacity)
          for (int a = 0; a < 3; ++a) ray_visited[3 * at + uint64_t(a)] = c[a];
        ++at;
      }
    }
    if (ray_offset) ray_offset[n_rays] = at;
    return 0;
  } catch (const std::exception& e) {
    g_error = e.what();
    return -1;
  }
}

}  // extern "C"
