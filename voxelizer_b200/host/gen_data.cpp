// vx_gen_data: drop-in for the reference's gen_data executable (src/gen_data.cpp:19-34):
//   vx_gen_data <config> <sequence-dirname> <output-voxel-dirname> [options]
// Same positional arguments, same output files, byte for byte.  Options (additions):
//   --device N            CUDA device (default 0)
//   --shard I/N           process slice I of N of the target list (multi-GPU: one process per GPU)
//   --frames-in-flight N  frame slots resident on the device
//   --first N --max N     sub-range of the (sharded) target list
//   --no-write            run the pipeline without writing files (timing)
//   --quiet
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <string>

#include "gen_data_pipeline.h"

int main(int argc, char** argv) {
  if (argc < 4) {
    std::cout << "Too few arguments" << std::endl;
    std::cout << "Usage: ./vx_gen_data <config> <sequence-dirname> <output-voxel-dirname> [--device N] [--shard I/N]"
                 " [--frames-in-flight N] [--first N] [--max N] [--no-write] [--quiet]"
              << std::endl;
    std::cout << "  <config>: config file, containing spatial extent of volume and resolution" << std::endl;
    std::cout << "  <sequence-dirname>: folder containing velodyne, labels, calib.txt, poses.txt" << std::endl;
    std::cout << "  <output-voxel-dirname>: output folder containing voxelized input and target volumes" << std::endl;
    return 1;
  }
  vxb::GenDataOptions opt;
  for (int i = 4; i < argc; ++i) {
    const std::string a = argv[i];
    auto next = [&]() -> const char* { return i + 1 < argc ? argv[++i] : ""; };
    if (a == "--device") opt.device = std::atoi(next());
    else if (a == "--shard") {
      unsigned si = 0, sn = 1;
      if (std::sscanf(next(), "%u/%u", &si, &sn) != 2 || sn == 0 || si >= sn) {
        std::cerr << "bad --shard, expected I/N" << std::endl;
        return 1;
      }
      opt.shardIndex = si;
      opt.shardCount = sn;
    } else if (a == "--frames-in-flight") opt.framesInFlight = uint32_t(std::atoi(next()));
    else if (a == "--first") opt.firstFrame = std::atoll(next());
    else if (a == "--max") opt.maxFrames = std::atoll(next());
    else if (a == "--no-write") opt.writeFiles = false;
    else if (a == "--quiet") opt.verbose = false;
    else {
      std::cerr << "unknown option " << a << std::endl;
      return 1;
    }
  }
  const vxb::Config config = vxb::parseConfiguration(argv[1]);
  vxb::GenDataStats st;
  std::string error;
  if (vxb::runGenData(config, argv[2], argv[3], opt, &st, &error) != 0) {
    std::cerr << "vx_gen_data: " << error << std::endl;
    return 2;
  }
  if (opt.verbose) {
    std::printf("[voxelizer_b200] frames %llu skipped %llu wall %.3fs (read %.3fs, gpu %.3fs, write-wait %.3fs)\n",
                (unsigned long long)st.frames, (unsigned long long)st.skipped, st.wall_s, st.read_s, st.gpu_s, st.write_s);
    std::printf("[voxelizer_b200] device ms: fill %.2f vote %.2f emit %.2f visibility %.2f pack %.2f | rays %llu steps %llu\n",
                st.stages.fill_ms, st.stages.vote_ms, st.stages.emit_ms, st.stages.visibility_ms, st.stages.pack_ms,
                (unsigned long long)st.stages.rays_cast, (unsigned long long)st.stages.dda_steps);
  }
  return 0;
}
