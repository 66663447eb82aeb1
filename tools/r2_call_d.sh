#!/bin/bash
# round-2 GPU call D: the middle shape under streaming load (909-frame steps back to back) and for one awaited step
mkdir -p gpurun_out
VX_DEV_VIS_MODE=3 timeout 300 python tools/dev_gpu_stream.py --steps 6 --warmup 3 --single --digest 2:0 > gpurun_out/r2e_mid_stream.log 2>&1
cat gpurun_out/r2e_mid_stream.log
