// The reference's gen_data main loop (src/gen_data.cpp:46-142) written against the kept host API --
// KittiReader::retrieve, VoxelGrid, fillVoxelGrid, saveVoxelGrid -- statement for statement.  It exists
// to show (and test) that a caller shaped like the reference's runs unchanged on the GPU backend; the
// production path is the batched pipeline in gen_data_pipeline.cpp (one frame at a time cannot fill a B200).
#include <cstdio>
#include <iostream>
#include <stdexcept>

#include "gen_data_pipeline.h"
#include "voxel_grid.h"

namespace vxb {

namespace {
float facadePoseDistance(const Matrix4f& A, const Matrix4f& B) {  // gen_data.cpp:15-17, Eigen norm order
  const float x = A(0, 3) - B(0, 3), y = A(1, 3) - B(1, 3), z = A(2, 3) - B(2, 3);
  return std::sqrt(x * x + (y * y + z * z));
}
}  // namespace

int runGenDataFacade(const Config& config, const std::string& sequences_dirname, const std::string& output_voxel_dirname, int device,
                     int64_t maxFrames, bool verbose, uint64_t* framesWritten, std::string* error) {
  try {
    KittiReader reader;
    reader.setVerbose(verbose);
    reader.initialize(sequences_dirname);

    reader.setNumPriorScans(int32_t(config.priorScans));
    reader.setNumPastScans(int32_t(config.pastScans));

    VoxelGrid priorGrid(device);
    VoxelGrid pastGrid(device);

    priorGrid.initialize(config.voxelSize, config.minExtent, config.maxExtent);
    pastGrid.initialize(config.voxelSize, config.minExtent, config.maxExtent);

    std::vector<Matrix4f> poses = reader.poses();

    uint32_t current = 0;
    uint64_t written = 0;

    while (current < reader.count()) {
      char outname[16];
      std::snprintf(outname, sizeof(outname), "%06u", current);

      std::vector<PointcloudPtr> priorPoints;
      std::vector<LabelsPtr> priorLabels;
      std::vector<PointcloudPtr> pastPoints;
      std::vector<LabelsPtr> pastLabels;

      reader.retrieve(int32_t(current), priorPoints, priorLabels, pastPoints, pastLabels);

      // ensure that labels are present, only then store data.
      uint32_t labelCount = 0;
      uint32_t pointCount = 0;
      for (uint32_t i = 0; i < pastLabels.size(); ++i) {
        pointCount += uint32_t(pastLabels[i]->size());
        for (uint32_t j = 0; j < pastLabels[i]->size(); ++j) {
          uint32_t label = (*pastLabels[i])[j];
          if (label > 0) labelCount += 1;
        }
      }
      float percentageLabeled = 100.0f * float(labelCount) / float(pointCount);

      priorGrid.clear();
      pastGrid.clear();

      if (percentageLabeled > 0.0f) {
        Matrix4f anchor_pose = priorPoints.back()->pose;

        fillVoxelGrid(anchor_pose, priorPoints, priorLabels, priorGrid, config);
        fillVoxelGrid(anchor_pose, priorPoints, priorLabels, pastGrid, config);
        fillVoxelGrid(anchor_pose, pastPoints, pastLabels, pastGrid, config);

        priorGrid.updateOcclusions();
        pastGrid.updateOcclusions();

        for (uint32_t i = 0; i < pastPoints.size(); ++i) {
          const Matrix4f ap = anchor_pose.inverse() * pastPoints[i]->pose;
          Vector3f endpoint(ap(0, 3), ap(1, 3), ap(2, 3));
          pastGrid.updateInvalid(endpoint);
        }

        // store grid
        saveVoxelGrid(priorGrid, output_voxel_dirname, outname, "input");
        saveVoxelGrid(pastGrid, output_voxel_dirname, outname, "target");
        ++written;
      } else {
        if (verbose) std::cout << "skipped." << std::endl;
      }
      if (maxFrames >= 0 && int64_t(written) >= maxFrames) break;

      // get index of next scan.
      float distance = 0.0f;
      uint32_t count = 0;
      while ((count < config.stride_num || distance < config.stride_distance || count == 0) && size_t(current) + 1 < poses.size()) {
        distance += facadePoseDistance(poses[current], poses[current + 1]);
        count += 1;
        current += 1;
      }
      if ((count < config.stride_num || distance < config.stride_distance || count == 0) && size_t(current) + 1 >= poses.size()) {
        break;  // no further scan can be extracted possible.
      }
    }
    if (framesWritten) *framesWritten = written;
    return 0;
  } catch (const std::exception& e) {
    if (error) *error = e.what();
    return -1;
  }
}

}  // namespace vxb
