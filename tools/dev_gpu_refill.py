"""Developer timing: the 909-frame job once, and k copies of it in ONE call (one visibility launch of k x 909 CTAs): how much
of a step is the tail of its slowest frames (CTAs of a longer launch start as resident ones retire)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import voxelizer_b200 as vx
from util import cfg_path

n_scans = 4541
CAP = 128000
cfg = vx.Config(cfg_path("semantic_kitti"))
params = vx.synth_params()
xyzr = np.empty((n_scans, CAP, 4), np.float32); lab = np.empty((n_scans, CAP), np.uint32)
counts = vx.synth_generate_many(params, 0, n_scans, xyzr.ctypes.data, lab.ctypes.data, CAP)
poses = vx.synth_velo_poses(params, n_scans).astype(np.float32)
be = vx.Backend.from_config(cfg)
for t in range(n_scans):
    be.upload_scan(t, xyzr[t, :counts[t]], lab[t, :counts[t]])
specs = []
for tgt in vx.plan_targets(cfg, n_scans, poses):
    pb, pe, qb, qe = vx.window(n_scans, cfg.pod.priorScans, cfg.pod.pastScans, int(tgt))
    a = poses[pe - 1]
    ap_p, _ = vx.frame_transforms(a, poses[pb:pe]); ap_q, ep = vx.frame_transforms(a, poses[qb:qe])
    specs.append(vx.FrameSpec(np.arange(pb, pe, dtype=np.uint32), ap_p, np.arange(qb, qe, dtype=np.uint32), ap_q, ep))
for k in (1, 2, 4):
    descs, keep = be.build_descs(specs * k, None)
    be.process_descs(descs, len(specs) * k)
    t = time.time()
    be.process_descs(descs, len(specs) * k)
    dt = time.time() - t
    print(f"{k} x {len(specs)} frames in one call: {dt * 1e3:.0f} ms, {k * len(specs) / dt:.1f} frames/s")
