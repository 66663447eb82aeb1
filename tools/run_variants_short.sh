# developer A/B of the fill / vote / emit stages: 1100 scans -> 220 frames (two fill groups), stage times only
for v in "$@"; do
  echo "=== $v"
  VX_B200_CUDA_LIB=$PWD/voxelizer_b200/lib/variants/libvx_$v.so python tools/dev_gpu_perf.py 1100 semantic_kitti 3 2>&1 | sed -n 2,3p
done
