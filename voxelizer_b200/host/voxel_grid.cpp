// VoxelGrid facade over the C ABI -- see voxel_grid.h.  No voxel arithmetic happens here: the grid
// geometry, fill, vote, visibility passes and packing all run in libvoxelizer_b200.so on the GPU.
#include "voxel_grid.h"

#include <cmath>
#include <cstring>
#include <fstream>
#include <limits>
#include <stdexcept>

#include "../csrc/vx_dda.h"
#include "gen_data_pipeline.h"

namespace vxb {

namespace {
[[noreturn]] void fail(vx_ctx* ctx, const std::string& what) {
  throw std::runtime_error(what + ": " + (ctx ? vx_last_error(ctx) : vx_last_global_error()));
}
}  // namespace

VoxelGrid::VoxelGrid(int device) : device_(device) {}

VoxelGrid::~VoxelGrid() {
  if (ctx_) vx_destroy(ctx_);
}

void VoxelGrid::requireInitialized() const {
  if (!initialized_) throw std::logic_error("VoxelGrid: initialize() has not been called");
}

void VoxelGrid::ensureContext(const Filter& f, bool occlusionLabels) const {
  if (ctx_ && ctxFilter_ == f && (ctxFill_ || !occlusionLabels)) return;
  if (ctx_) {
    vx_destroy(ctx_);
    ctx_ = nullptr;
  }
  vx_grid_desc gd;
  gd.resolution = resolution_;
  for (int a = 0; a < 3; ++a) {
    gd.min_extent[a] = min_[a];
    gd.max_extent[a] = max_[a];
  }
  vx_filter_desc fd;
  fd.min_range = f.minRange;
  fd.max_range = f.maxRange;
  fd.hidecar = f.hidecar ? 1 : 0;
  fd.filtered_labels = f.filtered.empty() ? nullptr : f.filtered.data();
  fd.n_filtered = uint32_t(f.filtered.size());
  fd.join_pairs = f.joins.empty() ? nullptr : f.joins.data();
  fd.n_join_pairs = uint32_t(f.joins.size() / 2);
  vx_options vo;
  std::memset(&vo, 0, sizeof(vo));
  vo.frames_in_flight = 1;  // one grid = one frame at a time
  vo.table_sets = 1;
  vo.debug = 1;             // keep the histogram for operator() / voxels()
  vo.occlusion_labels = occlusionLabels ? 1u : 0u;
  if (vx_create(device_, &gd, &fd, &vo, &ctx_) != VX_OK) fail(nullptr, "vx_create");
  ctxFilter_ = f;
  ctxFill_ = occlusionLabels;
}

void VoxelGrid::initialize(float resolution, const Vector4f& min, const Vector4f& max) {
  clear();
  resolution_ = resolution;
  min_ = min;
  max_ = max;
  initialized_ = true;
  // sizes and offset come from the library (VoxelGrid.cpp:8-18 arithmetic lives in vx_create)
  Filter none{0.0f, std::numeric_limits<float>::infinity(), false, {}, {}};
  if (ctx_) {
    vx_destroy(ctx_);
    ctx_ = nullptr;
  }
  ensureContext(none);
  vx_grid_info gi;
  if (vx_grid_info_get(ctx_, &gi) != VX_OK) fail(ctx_, "vx_grid_info_get");
  sizex_ = gi.size[0];
  sizey_ = gi.size[1];
  sizez_ = gi.size[2];
  offset_ = Vector4f(gi.offset[0], gi.offset[1], gi.offset[2], 1.0f);
}

void VoxelGrid::clear() {
  scans_.clear();
  mode_ = kEmpty;
  occlusions_ = false;
  fillLabels_ = false;
  endpoints_.clear();
  dirty_ = true;
  voxelsValid_ = false;
}

void VoxelGrid::insert(const Vector4f& p, uint32_t label) {
  requireInitialized();
  if (mode_ == kScans) throw std::logic_error("VoxelGrid: insert() and fillVoxelGrid() cannot be mixed between two clear() calls");
  if (occlusions_ || !endpoints_.empty()) throw std::logic_error("VoxelGrid: insert() after updateOcclusions()/updateInvalid() needs a clear()");
  if (mode_ == kEmpty) {
    mode_ = kRaw;
    scans_.emplace_back();
    scans_[0].ap = Matrix4f::Identity();
    filter_ = Filter{0.0f, std::numeric_limits<float>::infinity(), false, {}, {}};
  }
  PendingScan& s = scans_[0];
  s.xyzr.push_back(p.x());
  s.xyzr.push_back(p.y());
  s.xyzr.push_back(p.z());
  s.xyzr.push_back(0.0f);
  s.labels.push_back(label);
  dirty_ = true;
  voxelsValid_ = false;
}

void VoxelGrid::addScan(const Matrix4f& ap, const PointcloudPtr& points, const LabelsPtr& labels, const Config& config) {
  requireInitialized();
  if (mode_ == kRaw) throw std::logic_error("VoxelGrid: insert() and fillVoxelGrid() cannot be mixed between two clear() calls");
  if (occlusions_ || !endpoints_.empty()) throw std::logic_error("VoxelGrid: fillVoxelGrid() after updateOcclusions()/updateInvalid() needs a clear()");
  if (points->size() != labels->size()) throw std::runtime_error("Inconsistent number of labels.");
  Filter f{config.minRange, config.maxRange, config.hidecar, config.filteredLabels, {}};
  for (const auto& join : config.joinedLabels)
    for (const uint32_t member : join.second) {
      f.joins.push_back(join.first);
      f.joins.push_back(member);
    }
  if (mode_ == kScans && !(f == filter_)) throw std::logic_error("VoxelGrid: fillVoxelGrid() called with different Configs on one grid");
  filter_ = f;
  mode_ = kScans;
  PendingScan s;
  s.ap = ap;
  s.xyzr = points->xyzr;
  s.labels = *labels;
  scans_.push_back(std::move(s));
  dirty_ = true;
  voxelsValid_ = false;
}

void VoxelGrid::updateOcclusions() {
  requireInitialized();
  occlusions_ = true;
  dirty_ = true;
}

void VoxelGrid::insertOcclusionLabels() {
  requireInitialized();
  if (fillLabels_) throw std::logic_error("VoxelGrid: insertOcclusionLabels() once per clear()");
  if (!endpoints_.empty()) throw std::logic_error("VoxelGrid: insertOcclusionLabels() has to come before updateInvalid()");
  occlusions_ = true;  // if (!occlusionsValid_) updateOcclusions();  (VoxelGrid.cpp:76)
  fillLabels_ = true;
  dirty_ = true;
  voxelsValid_ = false;
}

void VoxelGrid::updateInvalid(const Vector3f& position) {
  requireInitialized();
  occlusions_ = true;  // the reference needs updateOcclusions() first (invalid_ = occlusions_, VoxelGrid.cpp:169)
  endpoints_.push_back(position);
  dirty_ = true;
}

void VoxelGrid::flush() const {
  requireInitialized();
  if (!dirty_) return;
  if (mode_ != kEmpty || fillLabels_) ensureContext(mode_ != kEmpty ? filter_ : ctxFilter_, fillLabels_);  // (an empty grid keeps initialize()'s context)
  if (vx_evict_all(ctx_) != VX_OK) fail(ctx_, "vx_evict_all");
  const uint32_t n = uint32_t(scans_.size());
  std::vector<uint32_t> ids(n);
  std::vector<float> ap(size_t(n) * 16);
  for (uint32_t s = 0; s < n; ++s) {
    ids[s] = s;
    std::memcpy(&ap[size_t(s) * 16], scans_[s].ap.m, 64);
    const uint32_t np = uint32_t(scans_[s].labels.size());
    if (vx_upload_scan(ctx_, s, scans_[s].xyzr.data(), scans_[s].labels.data(), np) != VX_OK) fail(ctx_, "vx_upload_scan");
  }
  std::vector<float> ep(endpoints_.size() * 3);
  for (size_t e = 0; e < endpoints_.size(); ++e)
    for (int a = 0; a < 3; ++a) ep[3 * e + size_t(a)] = endpoints_[e][a];
  const size_t nvox = num_elements();
  bin_.assign(nvox / 8, 0);
  occludedImg_.assign(nvox / 8, 0);
  invalidImg_.assign(nvox / 8, 0);
  label_.assign(nvox, 0);
  vx_frame_desc d;
  std::memset(&d, 0, sizeof(d));
  // the scans go in as PRIOR scans: they fill both device grids, so `.bin` (input mode) and
  // `.label/.occluded/.invalid` (target mode) all describe THIS grid
  d.n_prior = n;
  d.prior_ids = ids.data();
  d.ap_prior = ap.data();
  d.n_endpoints = uint32_t(endpoints_.size());
  d.endpoints = ep.data();
  d.out_bin = bin_.data();
  d.out_label = label_.data();
  d.out_occluded = occludedImg_.data();
  d.out_invalid = invalidImg_.data();
  if (fillLabels_) d.flags = VX_FRAME_INSERT_OCCLUSION_LABELS | VX_FRAME_INPUT_IS_TARGET;  // this grid is its own input grid
  if (vx_process_frames(ctx_, &d, 1) != VX_OK) fail(ctx_, "vx_process_frames");
  uint64_t count = 0;
  occludedFlag_.assign(nvox, 0);
  invalidFlag_.assign(nvox, 0);
  if (vx_debug_dump(ctx_, 0, VX_DUMP_OCCLUDED, occludedFlag_.data(), &count) != VX_OK) fail(ctx_, "vx_debug_dump");
  if (vx_debug_dump(ctx_, 0, VX_DUMP_INVALID, invalidFlag_.data(), &count) != VX_OK) fail(ctx_, "vx_debug_dump");
  dirty_ = false;
  voxelsValid_ = false;
}

const std::vector<VoxelGrid::Voxel>& VoxelGrid::voxels() const {
  flush();
  if (!voxelsValid_) {
    voxels_.assign(num_elements(), Voxel());
    uint64_t count = 0;
    if (vx_debug_dump(ctx_, 0, VX_DUMP_HIST_TARGET, nullptr, &count) != VX_OK) fail(ctx_, "vx_debug_dump");
    std::vector<uint32_t> triples(size_t(count) * 3);
    if (count && vx_debug_dump(ctx_, 0, VX_DUMP_HIST_TARGET, triples.data(), &count) != VX_OK) fail(ctx_, "vx_debug_dump");
    for (uint64_t t = 0; t < count; ++t) {
      const uint32_t L = triples[3 * t], label = triples[3 * t + 1], c = triples[3 * t + 2];
      const uint32_t k = L % sizez_, ij = L / sizez_;
      const uint32_t j = ij % sizey_, i = ij / sizey_;
      Voxel& v = voxels_[size_t(index(int32_t(i), int32_t(j), int32_t(k)))];
      v.labels[label] += c;
      v.count += c;
    }
    if (fillLabels_) {  // filled voxels hold a copy of their source voxel (VoxelGrid.cpp:90-91); sources are never filled themselves
      std::vector<uint32_t> src(num_elements());
      if (vx_debug_dump(ctx_, 0, VX_DUMP_FILL_SOURCE, src.data(), &count) != VX_OK) fail(ctx_, "vx_debug_dump");
      auto internal = [&](uint32_t L) {
        const uint32_t k = L % sizez_, ij = L / sizez_;
        return size_t(index(int32_t(ij / sizey_), int32_t(ij % sizey_), int32_t(k)));
      };
      for (uint32_t L = 0; L < num_elements(); ++L)
        if (src[L] != 0xFFFFFFFFu) voxels_[internal(L)] = voxels_[internal(src[L])];
    }
    voxelsValid_ = true;
  }
  return voxels_;
}

int32_t VoxelGrid::occludedBy(int32_t i, int32_t j, int32_t k, const Vector3f& endpoint, std::vector<Vector3i>* visited) const {
  flush();
  const float e[3] = {endpoint[0], endpoint[1], endpoint[2]};
  // a ray visits at most one voxel per step and every step moves one axis one way: Sx + Sy + Sz bounds the list
  const uint32_t capacity = visited ? sizex_ + sizey_ + sizez_ + 1u : 0u;
  std::vector<int32_t> ijk(size_t(capacity) * 3);
  uint32_t n = 0;
  int32_t occluder = -1;
  if (vx_trace_ray(ctx_, 0, i, j, k, e, capacity ? ijk.data() : nullptr, capacity, &n, &occluder) != VX_OK) fail(ctx_, "vx_trace_ray");
  if (visited) {
    visited->clear();
    for (uint32_t q = 0; q < n && q < capacity; ++q) visited->push_back(Vector3i(ijk[3 * q], ijk[3 * q + 1], ijk[3 * q + 2]));
  }
  return occluder;
}

bool VoxelGrid::isOccluded(int32_t i, int32_t j, int32_t k) const {
  flush();
  return occludedFlag_[(size_t(i) * sizey_ + size_t(j)) * sizez_ + size_t(k)] != 0;
}

bool VoxelGrid::isFree(int32_t i, int32_t j, int32_t k) const { return !isOccluded(i, j, k); }

bool VoxelGrid::isInvalid(int32_t i, int32_t j, int32_t k) const {
  flush();
  return invalidFlag_[(size_t(i) * sizey_ + size_t(j)) * sizez_ + size_t(k)] != 0;
}

Vector3f VoxelGrid::voxel2position(int32_t i, int32_t j, int32_t k) const {
  return Vector3f(vx_voxel_centre(offset_[0], resolution_, i), vx_voxel_centre(offset_[1], resolution_, j),
                  vx_voxel_centre(offset_[2], resolution_, k));
}

Vector3i VoxelGrid::position2voxel(const Vector3f& pos) const {
  return Vector3i(vx_trunc_i32((pos[0] - offset_[0]) / resolution_), vx_trunc_i32((pos[1] - offset_[1]) / resolution_),
                  vx_trunc_i32((pos[2] - offset_[2]) / resolution_));
}

const std::vector<uint8_t>& VoxelGrid::imageBin() const {
  flush();
  return bin_;
}
const std::vector<uint16_t>& VoxelGrid::imageLabel() const {
  flush();
  return label_;
}
const std::vector<uint8_t>& VoxelGrid::imageOccluded() const {
  flush();
  return occludedImg_;
}
const std::vector<uint8_t>& VoxelGrid::imageInvalid() const {
  flush();
  return invalidImg_;
}

void fillVoxelGrid(const Matrix4f& anchor_pose, const std::vector<PointcloudPtr>& points, const std::vector<LabelsPtr>& labels,
                   VoxelGrid& grid, const Config& config) {
  if (points.size() != labels.size()) throw std::runtime_error("Inconsistent number of labels.");
  for (size_t t = 0; t < points.size(); ++t) {
    float ap16[16];
    frameTransforms(anchor_pose, &points[t]->pose, 1, ap16, nullptr);  // ap = anchor_pose.inverse() * pose (voxelize_utils.cpp:183)
    Matrix4f ap;
    std::memcpy(ap.m, ap16, 64);
    grid.addScan(ap, points[t], labels[t], config);
  }
}

void saveVoxelGrid(const VoxelGrid& grid, const std::string& directory, const std::string& basename, const std::string& mode) {
  auto write = [&](const std::string& ext, const void* data, size_t bytes) {
    std::ofstream out(directory + "/" + basename + ext, std::ios::binary);
    out.write(static_cast<const char*>(data), std::streamsize(bytes));
  };
  if (mode == "target") {  // voxelize_utils.cpp:257-283
    write(".label", grid.imageLabel().data(), grid.imageLabel().size() * 2);
    write(".occluded", grid.imageOccluded().data(), grid.imageOccluded().size());
    write(".invalid", grid.imageInvalid().data(), grid.imageInvalid().size());
  } else {  // voxelize_utils.cpp:285-294
    write(".bin", grid.imageBin().data(), grid.imageBin().size());
  }
}

}  // namespace vxb
