// VoxelGrid / fillVoxelGrid / saveVoxelGrid: the reference's per-frame host API
// (src/data/VoxelGrid.h:17-135, src/data/voxelize_utils.h:33-48) kept by name, argument meaning and
// observable results on top of the C ABI (include/voxelizer_b200.h), for callers shaped like
// gen_data.cpp:89-121 and Mainframe::buildVoxelGrids (Mainframe.cpp:279-332):
//
//     grid.clear();
//     fillVoxelGrid(anchor_pose, points, labels, grid, config);
//     grid.updateOcclusions();
//     for (...) grid.updateInvalid(endpoint);
//     saveVoxelGrid(grid, dir, name, "target");   /  grid.isOccluded(i,j,k) ...
//
// The work is DEFERRED: fill / update calls record what to do, the first query (isOccluded, isFree,
// isInvalid, operator(), voxels(), saveVoxelGrid) runs ONE frame on the GPU (fill -> vote -> emit ->
// occlusion pass -> one invalid pass per recorded endpoint -> pack) and caches the results.  There is
// no CPU implementation behind it: without a CUDA device every query throws.
//
// Differences from the reference, all stated:
//   * insert() after updateOcclusions()/updateInvalid() without a clear() throws std::logic_error (the
//     reference would let later passes see the new points; no caller does that).
//   * insert() (raw, already transformed points) and fillVoxelGrid() (scan + pose + Config filters) cannot
//     be mixed on one grid between two clear() calls: std::logic_error.
//   * insertOcclusionLabels() (dead code in the reference: both call sites are commented out, gen_data.cpp:108-109,
//     Mainframe.cpp:303-304) is offered in the order those call sites use -- after the fills, before any updateInvalid(),
//     once per clear(); anything else throws std::logic_error.
//   * occludedBy() is the PUBLIC method (VoxelGrid.h:89-90): one ray on the device over the grid's occupancy, same
//     return value and visited list; its memo side effect is not observable through the public API (both passes
//     reset occludedBy_ before reading it) and does not exist here.
#ifndef VOXELIZER_B200_HOST_VOXEL_GRID_H_
#define VOXELIZER_B200_HOST_VOXEL_GRID_H_

#include <cstdint>
#include <map>
#include <string>
#include <vector>

#include "config.h"
#include "kitti_reader.h"
#include "voxelizer_b200.h"

namespace vxb {

struct Vector3i {
  int32_t v[3];
  Vector3i() : v{0, 0, 0} {}
  Vector3i(int32_t a, int32_t b, int32_t c) : v{a, b, c} {}
  int32_t& operator[](int i) { return v[i]; }
  const int32_t& operator[](int i) const { return v[i]; }
};

class VoxelGrid {
 public:
  class Voxel {
   public:
    std::map<uint32_t, uint32_t> labels;
    uint32_t count{0};
  };

  explicit VoxelGrid(int device = 0);
  ~VoxelGrid();
  VoxelGrid(const VoxelGrid&) = delete;
  VoxelGrid& operator=(const VoxelGrid&) = delete;

  /** set parameters of voxel grid and compute internal offsets (VoxelGrid.cpp:5-28) */
  void initialize(float resolution, const Vector4f& min, const Vector4f& max);

  const Voxel& operator()(uint32_t i, uint32_t j, uint32_t k) const { return voxels()[size_t(index(int32_t(i), int32_t(j), int32_t(k)))]; }
  void clear();
  /** add label for an already transformed point (VoxelGrid.cpp:45-63) */
  void insert(const Vector4f& p, uint32_t label);
  const std::vector<Voxel>& voxels() const;

  const Vector4f& offset() const { return offset_; }
  uint32_t num_elements() const { return sizex_ * sizey_ * sizez_; }
  uint32_t size(uint32_t dim) const { return dim == 0 ? sizex_ : (dim == 1 ? sizey_ : sizez_); }
  float resolution() const { return resolution_; }

  void updateOcclusions();
  void updateInvalid(const Vector3f& position);
  /** fill occluded voxels with the labels of the occupied voxel above them (VoxelGrid.cpp:75-96) */
  void insertOcclusionLabels();

  /** ray from voxel (i,j,k) towards `endpoint`: index() of the occupied voxel that stops it or -1; `visited` receives
   *  every voxel the ray looks at, the occluder included (VoxelGrid.h:89-90, VoxelGrid.cpp:235-316) */
  int32_t occludedBy(int32_t i, int32_t j, int32_t k, const Vector3f& endpoint = Vector3f(0.0f, 0.0f, 0.0f),
                     std::vector<Vector3i>* visited = nullptr) const;

  bool isOccluded(int32_t i, int32_t j, int32_t k) const;
  bool isFree(int32_t i, int32_t j, int32_t k) const;
  bool isInvalid(int32_t i, int32_t j, int32_t k) const;

  Vector3f voxel2position(int32_t i, int32_t j, int32_t k) const;
  Vector3i position2voxel(const Vector3f& pos) const;
  int32_t index(int32_t i, int32_t j, int32_t k) const { return i + j * int32_t(sizex_) + k * int32_t(sizex_ * sizey_); }

  // ---- used by fillVoxelGrid / saveVoxelGrid below
  void addScan(const Matrix4f& ap, const PointcloudPtr& points, const LabelsPtr& labels, const Config& config);
  /** the four file images of saveVoxelGrid (voxelize_utils.cpp:218-296), computed on the device */
  const std::vector<uint8_t>& imageBin() const;
  const std::vector<uint16_t>& imageLabel() const;
  const std::vector<uint8_t>& imageOccluded() const;
  const std::vector<uint8_t>& imageInvalid() const;

 private:
  struct Filter {
    float minRange, maxRange;
    bool hidecar;
    std::vector<uint32_t> filtered;
    std::vector<uint32_t> joins;  // (key, member) pairs
    bool operator==(const Filter& o) const {
      return minRange == o.minRange && maxRange == o.maxRange && hidecar == o.hidecar && filtered == o.filtered && joins == o.joins;
    }
  };
  struct PendingScan {
    Matrix4f ap;
    std::vector<float> xyzr;
    std::vector<uint32_t> labels;
  };
  void ensureContext(const Filter& f, bool occlusionLabels = false) const;
  void flush() const;
  void requireInitialized() const;

  int device_;
  bool initialized_{false};
  float resolution_{0};
  Vector4f min_, max_, offset_;
  uint32_t sizex_{0}, sizey_{0}, sizez_{0};

  enum Mode { kEmpty, kRaw, kScans };
  Mode mode_{kEmpty};
  Filter filter_{};
  std::vector<PendingScan> scans_;  // kScans: one per fillVoxelGrid scan; kRaw: scans_[0] collects insert() points
  bool occlusions_{false};
  bool fillLabels_{false};  // insertOcclusionLabels() was called
  std::vector<Vector3f> endpoints_;

  // device side + cached results
  mutable vx_ctx* ctx_{nullptr};
  mutable Filter ctxFilter_{};
  mutable bool ctxFill_{false};
  mutable bool dirty_{true};
  mutable std::vector<uint8_t> bin_, occludedImg_, invalidImg_, occludedFlag_, invalidFlag_;
  mutable std::vector<uint16_t> label_;
  mutable std::vector<Voxel> voxels_;
  mutable bool voxelsValid_{false};
};

/** use given anchor_pose to insert all point clouds into the given VoxelGrid (voxelize_utils.cpp:173-203) */
void fillVoxelGrid(const Matrix4f& anchor_pose, const std::vector<PointcloudPtr>& points, const std::vector<LabelsPtr>& labels,
                   VoxelGrid& grid, const Config& config);

/** write <directory>/<basename>.bin ("input") or .label/.occluded/.invalid ("target") (voxelize_utils.cpp:218-296) */
void saveVoxelGrid(const VoxelGrid& grid, const std::string& directory, const std::string& basename, const std::string& mode = "input");

}  // namespace vxb

#endif
