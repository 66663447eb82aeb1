# developer A/B: each variant library through tools/dev_gpu_perf.py (909 frames of configs[1])
for v in "$@"; do
  echo "=== $v"
  VX_B200_CUDA_LIB=$PWD/voxelizer_b200/lib/variants/libvx_$v.so python tools/dev_gpu_perf.py 4541 semantic_kitti 2 2>&1 | tail -7
done
