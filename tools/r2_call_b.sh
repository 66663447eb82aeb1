#!/bin/bash
# round-2 GPU call B: timeline of the host-buffer path with 3 / 1 / 4 steps in flight
mkdir -p gpurun_out
rm -f gpurun_out/r2c_*
export VX_B200_TRACE=gpurun_out/r2c_trace VX_B200_FRAME_STATS=gpurun_out/r2c_frames
timeout 300 python tools/dev_gpu_e2e_stream.py 5 3:3:1 1:2:0 4:4:1 > gpurun_out/r2c_e2e_stream.log 2>&1
cat gpurun_out/r2c_e2e_stream.log
ls -la gpurun_out/r2c_*
