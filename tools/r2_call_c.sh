#!/bin/bash
# round-2 GPU call C: the new tests, the host-buffer path with 2-D copies (timeline), the middle traversal shape on a 568-frame job
mkdir -p gpurun_out
rm -f gpurun_out/r2d_*
timeout 300 python -m pytest tests/test_gpu_parity.py -q -x -k "ring or several_jobs or async_submit or fill_groups or insert_occlusion" > gpurun_out/r2d_tests.log 2>&1; echo "pytest rc $?" >> gpurun_out/r2d_tests.log
tail -4 gpurun_out/r2d_tests.log
export VX_B200_TRACE=gpurun_out/r2d_trace VX_B200_FRAME_STATS=gpurun_out/r2d_frames
timeout 300 python tools/dev_gpu_e2e_stream.py 5 3:3:1 2:2:0 > gpurun_out/r2d_e2e_stream.log 2>&1
cat gpurun_out/r2d_e2e_stream.log
unset VX_B200_TRACE VX_B200_FRAME_STATS
for m in 1 3; do echo "=== 568 dense frames, vis_mode $m"; VX_DEV_VIS_MODE=$m timeout 200 python tools/dev_gpu_perf.py 568 semantic_kitti_dense 2 2>&1 | sed -n 1,2p; done > gpurun_out/r2d_mid.log 2>&1
cat gpurun_out/r2d_mid.log
