#!/bin/bash
# round-2 GPU call A: full GPU suite with the 3-lane pipeline, then the lane / priority / CTA-shape A/B, then a short bench with e2e
mkdir -p gpurun_out
V=$PWD/voxelizer_b200/lib/variants
timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/r2b_gputest.log 2>&1; echo "pytest rc $?" >> gpurun_out/r2b_gputest.log
tail -3 gpurun_out/r2b_gputest.log
{
timeout 300 python tools/dev_gpu_stream.py --steps 8 --warmup 4 --digest --single 2:0 3:0 3:1 4:1
VX_B200_CUDA_LIB=$V/libvx_c8.so timeout 200 python tools/dev_gpu_stream.py --steps 8 --warmup 4 --digest 3:1 4:1
VX_B200_CUDA_LIB=$V/libvx_t64b.so timeout 200 python tools/dev_gpu_stream.py --steps 8 --warmup 5 --digest 4:1
VX_B200_CUDA_LIB=$V/libvx_t64.so timeout 200 python tools/dev_gpu_stream.py --steps 8 --warmup 5 --digest 4:1
} > gpurun_out/r2b_stream.log 2>&1
cat gpurun_out/r2b_stream.log
timeout 400 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/r2b_bench_short.json 2> gpurun_out/r2b_bench_short.err
echo "bench rc $?"; tail -c 600 gpurun_out/r2b_bench_short.err
python - <<'P'
import json
try:
    d = json.loads(open('gpurun_out/r2b_bench_short.json').read().strip().splitlines()[-1])
    print('value', d['value'], 'single', d['single_step']['value'], 'e2e', d['e2e'], 'frac', d['roofline']['frac'], d['roofline']['insert_vote_pack'], d['roofline']['overlapped'])
except Exception as e:
    print('no bench line', e)
P
