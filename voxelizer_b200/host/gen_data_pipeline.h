// The gen_data frame loop (src/gen_data.cpp:19-145) re-organised for a GPU: the target list and the
// windows are planned on the host exactly as the reference walks them, the frames are then pushed
// through the C ABI (include/voxelizer_b200.h) in batches, and the four files per frame are written by
// a small pool of writer threads while the next batch is on the device.  Scans are read by a pool of
// reader threads through a small pinned staging ring; outputs land in pageable memory.
#ifndef VOXELIZER_B200_HOST_GEN_DATA_PIPELINE_H_
#define VOXELIZER_B200_HOST_GEN_DATA_PIPELINE_H_

#include <cstdint>
#include <string>
#include <vector>

#include "config.h"
#include "kitti_reader.h"
#include "voxelizer_b200.h"

namespace vxb {

struct GenDataOptions {
  int device = 0;
  uint32_t framesInFlight = 0;  // 0 = library default
  uint32_t shardIndex = 0;      // this process handles slice shardIndex of shardCount of the target list
  uint32_t shardCount = 1;
  int64_t firstFrame = 0;       // position in the (sharded) target list to start at
  int64_t maxFrames = -1;       // -1 = all
  bool writeFiles = true;
  bool verbose = true;
  int ioThreads = 8;    // file writers
  int readThreads = 8;  // file readers (disk -> pinned staging ring -> device)
};

struct GenDataStats {
  uint64_t frames = 0;   // frames written
  uint64_t skipped = 0;  // frames that failed the "labelled" gate (gen_data.cpp:92,124-126)
  double wall_s = 0;
  double read_s = 0;     // disk -> pinned host -> device
  double gpu_s = 0;      // vx_process_frames
  double write_s = 0;    // file writes not hidden behind GPU work
  vx_stage_times stages{};
};

// Target scan indices in processing order: a pure function of (number of .bin files, poses, stride
// settings) -- the "skipped." gate does not influence the walk (gen_data.cpp:62,129-141).
std::vector<uint32_t> planTargets(uint32_t scanCount, const std::vector<Matrix4f>& poses, const Config& config);

// ap = anchor^-1 * pose for each pose (voxelize_utils.cpp:183); endpoints = translation column of the
// same product (gen_data.cpp:114).  ap16: n x 16 column-major, endpoints: n x 3 (may be null).
void frameTransforms(const Matrix4f& anchor, const Matrix4f* poses, uint32_t n, float* ap16, float* endpoints);

// grid / filter descriptors of the C ABI from a Config (join_pairs storage is owned by `scratch`).
void describe(const Config& config, vx_grid_desc& grid, vx_filter_desc& filter, std::vector<uint32_t>& scratch);

int runGenData(const Config& config, const std::string& sequenceDir, const std::string& outputDir,
               const GenDataOptions& options, GenDataStats* stats, std::string* error);

}  // namespace vxb

#endif
