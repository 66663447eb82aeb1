// KittiReader: KITTI-format sequence access with the reference's semantics
// (src/widget/KittiReader.{h,cpp}, src/data/kitti_utils.cpp:29-58,71-105) and without Qt:
// velodyne/*.bin listed sorted by name, label file = name up to the FIRST '.' + ".label", missing
// label dir / files are CREATED zero-filled, poses = Tr^-1 * P * Tr, labels masked with 0xFFFF,
// windows prior = [max(0, idx - prior), idx], past = [min(N-1, idx+1), min(N, idx + past + 1)).
#ifndef VOXELIZER_B200_HOST_KITTI_READER_H_
#define VOXELIZER_B200_HOST_KITTI_READER_H_

#include <cstdint>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "pose_math.h"

namespace vxb {

struct Point3f {
  float x, y, z;
};

// One scan.  Points are kept in the on-disk layout (x, y, z, remission) -- also the device layout.
class Laserscan {
 public:
  void clear() { xyzr.clear(); }
  uint32_t size() const { return static_cast<uint32_t>(xyzr.size() / 4); }
  Point3f point(uint32_t i) const { return Point3f{xyzr[4 * size_t(i)], xyzr[4 * size_t(i) + 1], xyzr[4 * size_t(i) + 2]}; }
  float remission(uint32_t i) const { return xyzr[4 * size_t(i) + 3]; }
  bool hasRemissions() const { return !xyzr.empty(); }

  Matrix4f pose;
  std::vector<float> xyzr;
};

typedef std::shared_ptr<Laserscan> PointcloudPtr;
typedef std::shared_ptr<std::vector<uint32_t>> LabelsPtr;

class KITTICalibration {
 public:
  void initialize(const std::string& filename);
  const Matrix4f& operator[](const std::string& name) const;
  bool exists(const std::string& name) const { return m_.count(name) != 0; }
  void clear() { m_.clear(); }

 private:
  std::map<std::string, Matrix4f> m_;
};

std::vector<Matrix4f> loadPoses(const std::string& filename);

class KittiReader {
 public:
  void initialize(const std::string& directory);

  uint32_t count() const { return static_cast<uint32_t>(velodyne_filenames_.size()); }
  void setNumPriorScans(int32_t count) { numPriorScans_ = count; }
  void setNumPastScans(int32_t count) { numPastScans_ = count; }

  /** points and labels of the window around idx; prints "<k> point clouds read." like the reference. */
  void retrieve(int32_t idx, std::vector<PointcloudPtr>& priorPoints, std::vector<LabelsPtr>& priorLabels,
                std::vector<PointcloudPtr>& pastPoints, std::vector<LabelsPtr>& pastLabels);

  std::vector<Matrix4f> poses() const { return poses_; }

  // ---- additions used by the batched GPU pipeline (no reference counterpart)
  /** scan index ranges [begin, end) of the windows retrieve(idx) would return. */
  void window(int32_t idx, int32_t& priorBegin, int32_t& priorEnd, int32_t& pastBegin, int32_t& pastEnd) const;
  /** reads scan t from disk (no cache); throws like retrieve() on a label-count mismatch. */
  void load(uint32_t t, Laserscan& scan, std::vector<uint32_t>& labels) const;
  /** the same straight into caller memory (e.g. a pinned staging slot of `capacity` points): raw labels (the
   *  device applies & 0xFFFF), `labelled` = labels > 0 after the mask.  false if the scan does not fit. */
  bool loadInto(uint32_t t, float* xyzr, uint32_t* labels, uint32_t capacity, uint32_t& points, uint32_t& labelled) const;
  const Matrix4f& pose(uint32_t t) const { return poses_.at(t); }
  const std::string& velodyneFilename(uint32_t t) const { return velodyne_filenames_[t]; }
  void setVerbose(bool v) { verbose_ = v; }

 protected:
  void readPoints(const std::string& filename, Laserscan& scan) const;
  void readLabels(const std::string& filename, std::vector<uint32_t>& labels) const;
  void readPoses(const std::string& filename, std::vector<Matrix4f>& poses);

  KITTICalibration calib_;
  std::vector<Matrix4f> poses_;
  std::vector<std::string> velodyne_filenames_;
  std::vector<std::string> label_filenames_;
  std::map<uint32_t, PointcloudPtr> pointsCache_;
  std::map<uint32_t, LabelsPtr> labelCache_;
  int32_t numPriorScans_{0};
  int32_t numPastScans_{10};
  bool verbose_{true};
};

}  // namespace vxb

#endif
