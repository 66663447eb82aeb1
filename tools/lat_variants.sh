for v in lat1024 lat768; do echo "=== $v"; VX_B200_CUDA_LIB=$PWD/voxelizer_b200/lib/variants/libvx_$v.so python tools/dev_gpu_latency.py 2>&1 | tail -1; VX_B200_CUDA_LIB=$PWD/voxelizer_b200/lib/variants/libvx_$v.so python tools/dev_gpu_perf.py 200 stress_512 1 2>&1 | sed -n 1,2p; done
echo "=== default"; python tools/dev_gpu_perf.py 200 stress_512 1 2>&1 | sed -n 1,2p
