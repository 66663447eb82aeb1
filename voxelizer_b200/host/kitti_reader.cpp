#include "kitti_reader.h"

#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>

#include <dirent.h>
#include <sys/stat.h>

#include <algorithm>
#include <fstream>
#include <iostream>
#include <stdexcept>

#include "config.h"  // trim / split / lexical_float

namespace vxb {

namespace {

bool isDirectory(const std::string& p) {
  struct stat st;
  return ::stat(p.c_str(), &st) == 0 && S_ISDIR(st.st_mode);
}

bool pathExists(const std::string& p) {
  struct stat st;
  return ::stat(p.c_str(), &st) == 0;
}

uint64_t streamSize(std::ifstream& in) {
  in.seekg(0, std::ios::end);
  const std::streamoff n = in.tellg();
  in.seekg(0, std::ios::beg);
  return n < 0 ? 0u : static_cast<uint64_t>(n);
}

// 12 row-major numbers -> top 3x4 of an identity matrix
Matrix4f fromRowMajor12(const std::vector<std::string>& entries) {
  Matrix4f m = Matrix4f::Identity();
  for (int i = 0; i < 12; ++i) m(i / 4, i % 4) = lexical_float(trim(entries[size_t(i)]));
  return m;
}

}  // namespace

void KITTICalibration::initialize(const std::string& filename) {
  m_.clear();
  std::ifstream in(filename.c_str());
  if (!in.is_open()) throw std::runtime_error("Unable to open calibration file: " + filename);
  std::string line;
  in.peek();
  while (!in.eof()) {
    std::getline(in, line);
    const std::vector<std::string> kv = split(line, ":");
    if (kv.size() == 2) {
      const std::vector<std::string> entries = split(trim(kv[1]), " ");
      if (entries.size() == 12) m_[trim(kv[0])] = fromRowMajor12(entries);
    }
    in.peek();
  }
}

const Matrix4f& KITTICalibration::operator[](const std::string& name) const {
  const auto it = m_.find(name);
  if (it == m_.end()) throw std::runtime_error("Unknown name for calibration matrix.");
  return it->second;
}

std::vector<Matrix4f> loadPoses(const std::string& filename) {
  std::ifstream in(filename.c_str());
  if (!in.is_open()) throw std::runtime_error("Unable to open pose file :" + filename);
  std::vector<Matrix4f> poses;
  std::string line;
  in.peek();
  while (in.good()) {
    std::getline(in, line);
    const std::vector<std::string> entries = split(line, " ");  // empty tokens count
    if (entries.size() >= 12) poses.push_back(fromRowMajor12(entries));
    in.peek();
  }
  return poses;
}

void KittiReader::initialize(const std::string& directory) {
  velodyne_filenames_.clear();
  label_filenames_.clear();
  pointsCache_.clear();
  labelCache_.clear();

  const std::string velodyne = directory + "/velodyne";
  if (!isDirectory(velodyne)) throw std::runtime_error("Missing velodyne files.");
  std::vector<std::string> names;
  if (DIR* d = ::opendir(velodyne.c_str())) {
    while (const dirent* e = ::readdir(d)) {
      const std::string name = e->d_name;
      if (name.empty() || name[0] == '.') continue;
      struct stat st;
      if (::stat((velodyne + "/" + name).c_str(), &st) == 0 && S_ISREG(st.st_mode)) names.push_back(name);
    }
    ::closedir(d);
  }
  std::sort(names.begin(), names.end());
  for (const std::string& n : names) velodyne_filenames_.push_back(velodyne + "/" + n);

  const std::string calib = directory + "/calib.txt";
  if (!pathExists(calib)) throw std::runtime_error("Missing calibration file: " + calib);
  calib_.initialize(calib);
  readPoses(directory + "/poses.txt", poses_);

  const std::string labels = directory + "/labels";
  if (!isDirectory(labels)) ::mkdir(labels.c_str(), 0777);
  for (size_t i = 0; i < names.size(); ++i) {
    std::ifstream in(velodyne_filenames_[i].c_str());
    const uint32_t numPoints = static_cast<uint32_t>(streamSize(in) / 16);
    in.close();
    const std::string labelFile = labels + "/" + names[i].substr(0, names[i].find('.')) + ".label";
    if (!pathExists(labelFile)) {  // the reference creates all-zero labels as a side effect
      std::ofstream out(labelFile.c_str());
      const std::vector<uint32_t> zeros(numPoints, 0);
      out.write(reinterpret_cast<const char*>(zeros.data()), std::streamsize(size_t(numPoints) * 4));
    }
    label_filenames_.push_back(labelFile);
  }
}

void KittiReader::readPoses(const std::string& filename, std::vector<Matrix4f>& poses) {
  poses = loadPoses(filename);
  const Matrix4f Tr = calib_["Tr"];
  const Matrix4f Tr_inv = Tr.inverse();
  for (Matrix4f& p : poses) p = (Tr_inv * p) * Tr;  // camera -> velodyne frame, left to right
}

void KittiReader::readPoints(const std::string& filename, Laserscan& scan) const {
  std::ifstream in(filename.c_str(), std::ios::binary);
  if (!in.is_open()) return;  // silently empty, as the reference
  scan.clear();
  const uint32_t numPoints = static_cast<uint32_t>(streamSize(in) / 16);
  scan.xyzr.resize(size_t(numPoints) * 4);
  in.read(reinterpret_cast<char*>(scan.xyzr.data()), std::streamsize(size_t(numPoints) * 16));
}

void KittiReader::readLabels(const std::string& filename, std::vector<uint32_t>& labels) const {
  std::ifstream in(filename.c_str(), std::ios::binary);
  if (!in.is_open()) {
    std::cerr << "Unable to open label file. " << std::endl;
    return;
  }
  labels.clear();
  const uint32_t n = static_cast<uint32_t>(streamSize(in) / 4);
  labels.resize(n);
  in.read(reinterpret_cast<char*>(labels.data()), std::streamsize(size_t(n) * 4));
  for (uint32_t& v : labels) v &= 0xFFFFu;  // drop the instance id
}

void KittiReader::load(uint32_t t, Laserscan& scan, std::vector<uint32_t>& labels) const {
  scan.clear();
  labels.clear();
  readPoints(velodyne_filenames_.at(t), scan);
  scan.pose = poses_.at(t);
  readLabels(label_filenames_.at(t), labels);
  if (scan.size() != labels.size()) {
    std::cout << "Filename: " << velodyne_filenames_[t] << std::endl;
    std::cout << "Filename: " << label_filenames_[t] << std::endl;
    std::cout << "num. points = " << scan.size() << " vs. num. labels = " << labels.size() << std::endl;
    throw std::runtime_error("Inconsistent number of labels.");
  }
}

namespace {

// whole file into `dst` (at most `capacity` bytes); returns the file size, or -1 if it cannot be opened.  A short
// read (EIO, or the file shrank between fstat and read) throws: the tail of `dst` still holds an older scan.
int64_t slurp(const std::string& filename, void* dst, size_t capacity) {
  const int fd = ::open(filename.c_str(), O_RDONLY);
  if (fd < 0) return -1;
  struct stat st;
  if (::fstat(fd, &st) != 0) {
    ::close(fd);
    return -1;
  }
  const size_t want = std::min<size_t>(size_t(st.st_size), capacity);
  size_t got = 0;
  while (got < want) {
    const ssize_t n = ::read(fd, static_cast<char*>(dst) + got, want - got);
    if (n <= 0) break;
    got += size_t(n);
  }
  ::close(fd);
  if (got != want) throw std::runtime_error("short read: " + filename);
  return int64_t(st.st_size);
}

}  // namespace

bool KittiReader::loadInto(uint32_t t, float* xyzr, uint32_t* labels, uint32_t capacity, uint32_t& points, uint32_t& labelled) const {
  const int64_t pointBytes = slurp(velodyne_filenames_.at(t), xyzr, size_t(capacity) * 16);
  const int64_t labelBytes = slurp(label_filenames_.at(t), labels, size_t(capacity) * 4);
  const uint32_t np = pointBytes < 0 ? 0u : uint32_t(pointBytes / 16);  // unreadable = empty, as readPoints / readLabels
  const uint32_t nl = labelBytes < 0 ? 0u : uint32_t(labelBytes / 4);
  if (labelBytes < 0) std::cerr << "Unable to open label file. " << std::endl;
  if (np != nl) {
    std::cout << "Filename: " << velodyne_filenames_[t] << std::endl;
    std::cout << "Filename: " << label_filenames_[t] << std::endl;
    std::cout << "num. points = " << np << " vs. num. labels = " << nl << std::endl;
    throw std::runtime_error("Inconsistent number of labels.");
  }
  if (np > capacity) return false;
  points = np;
  uint32_t count = 0;
  for (uint32_t i = 0; i < np; ++i) count += (labels[i] & 0xFFFFu) > 0 ? 1u : 0u;
  labelled = count;
  return true;
}

void KittiReader::window(int32_t idx, int32_t& priorBegin, int32_t& priorEnd, int32_t& pastBegin, int32_t& pastEnd) const {
  const int32_t n = static_cast<int32_t>(velodyne_filenames_.size());
  priorBegin = std::max<int32_t>(0, idx - numPriorScans_);
  priorEnd = idx + 1;
  pastBegin = std::min<int32_t>(n - 1, idx + 1);  // == idx for the last scan: it is inserted twice
  pastEnd = std::min<int32_t>(n, idx + numPastScans_ + 1);
  if (pastEnd < pastBegin) pastEnd = pastBegin;
}

void KittiReader::retrieve(int32_t idx, std::vector<PointcloudPtr>& priorPoints, std::vector<LabelsPtr>& priorLabels,
                           std::vector<PointcloudPtr>& pastPoints, std::vector<LabelsPtr>& pastLabels) {
  priorPoints.clear();
  priorLabels.clear();
  pastPoints.clear();
  pastLabels.clear();
  if (idx < 0 || idx >= static_cast<int32_t>(velodyne_filenames_.size())) return;

  int32_t pb, pe, qb, qe;
  window(idx, pb, pe, qb, qe);
  uint32_t scansRead = 0;
  std::map<uint32_t, PointcloudPtr> keepPoints;
  std::map<uint32_t, LabelsPtr> keepLabels;
  auto fetch = [&](int32_t t, std::vector<PointcloudPtr>& pts, std::vector<LabelsPtr>& labs) {
    auto hit = pointsCache_.find(uint32_t(t));
    if (hit == pointsCache_.end()) {
      ++scansRead;
      PointcloudPtr scan(new Laserscan);
      LabelsPtr labels(new std::vector<uint32_t>());
      load(uint32_t(t), *scan, *labels);
      pointsCache_[uint32_t(t)] = scan;
      labelCache_[uint32_t(t)] = labels;
    }
    pts.push_back(pointsCache_[uint32_t(t)]);
    labs.push_back(labelCache_[uint32_t(t)]);
    keepPoints[uint32_t(t)] = pts.back();
    keepLabels[uint32_t(t)] = labs.back();
  };
  for (int32_t t = pb; t < pe; ++t) fetch(t, priorPoints, priorLabels);
  for (int32_t t = qb; t < qe; ++t) fetch(t, pastPoints, pastLabels);
  if (verbose_) std::cout << scansRead << " point clouds read." << std::endl;
  pointsCache_.swap(keepPoints);  // everything outside the new window is evicted
  labelCache_.swap(keepLabels);
}

}  // namespace vxb
