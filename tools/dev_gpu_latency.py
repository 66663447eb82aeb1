"""Developer timing: ONE full-window frame (1 + 70 scans, semantic_kitti.cfg) through the C ABI in both shapes of the
traversal kernel -- the time a GUI rebuild (Mainframe::buildVoxelGrids) waits for."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import voxelizer_b200 as vx
from util import cfg_path

cfg_name = sys.argv[1] if len(sys.argv) > 1 else "semantic_kitti"
CAP = 128000
cfg = vx.Config(cfg_path(cfg_name))
n = cfg.pod.pastScans + 1
params = vx.synth_params()
xyzr = np.empty((n, CAP, 4), np.float32); lab = np.empty((n, CAP), np.uint32)
counts = vx.synth_generate_many(params, 0, n, xyzr.ctypes.data, lab.ctypes.data, CAP)
poses = vx.synth_velo_poses(params, n).astype(np.float32)
for mode, name in ((1, "throughput shape (128 threads)"), (2, "latency shape (512 threads)")):
    be = vx.Backend.from_config(cfg, frames_in_flight=1, vis_mode=mode)
    for t in range(n):
        be.upload_scan(t, xyzr[t, :counts[t]], lab[t, :counts[t]])
    ap_p, _ = vx.frame_transforms(poses[0], poses[0:1]); ap_q, ep = vx.frame_transforms(poses[0], poses[1:n])
    spec = vx.FrameSpec(np.arange(0, 1, dtype=np.uint32), ap_p, np.arange(1, n, dtype=np.uint32), ap_q, ep)
    descs, keep = be.build_descs([spec], None)
    be.process_descs(descs, 1)
    be.reset_stage_times()
    t0 = time.time()
    for _ in range(3):
        be.process_descs(descs, 1)
    dt = (time.time() - t0) / 3
    st = be.stage_times()
    print(f"{cfg_name} one frame, {n - 1} passes, {name}: {dt * 1e3:.0f} ms (visibility {st['visibility_ms'] / 3:.0f} ms, waves {st['vis_waves'] / 3:.0f})")
    be.close()
