"""Developer A/B in ONE process: the same N-frame job of a config in several traversal shapes (vx_options.vis_mode).
    python tools/dev_gpu_modes.py n_scans cfg_name mode [mode ...]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import voxelizer_b200 as vx
from util import cfg_path

n_scans, cfg_name, modes = int(sys.argv[1]), sys.argv[2], [int(m) for m in sys.argv[3:]]
CAP = 128000
cfg = vx.Config(cfg_path(cfg_name))
params = vx.synth_params()
px, pl = vx.PinnedBuffer(n_scans * CAP * 16), vx.PinnedBuffer(n_scans * CAP * 4)
counts = vx.synth_generate_many(params, 0, n_scans, px.ptr, pl.ptr, CAP, threads=os.cpu_count() or 8)
poses = vx.synth_velo_poses(params, n_scans).astype(np.float32)
specs = []
for tgt in vx.plan_targets(cfg, n_scans, poses):
    pb, pe, qb, qe = vx.window(n_scans, cfg.pod.priorScans, cfg.pod.pastScans, int(tgt))
    a = poses[pe - 1]
    ap_p, _ = vx.frame_transforms(a, poses[pb:pe]); ap_q, ep = vx.frame_transforms(a, poses[qb:qe])
    specs.append(vx.FrameSpec(np.arange(pb, pe, dtype=np.uint32), ap_p, np.arange(qb, qe, dtype=np.uint32), ap_q, ep))
for mode in modes:
    be = vx.Backend.from_config(cfg, vis_mode=mode)
    for t in range(n_scans):
        be.upload_scan_ptr(t, px.ptr + t * CAP * 16, pl.ptr + t * CAP * 4, int(counts[t]))
    descs, keep = be.build_descs(specs, None)
    be.process_descs(descs, len(specs))
    t0 = time.time()
    be.process_descs(descs, len(specs))
    dt = time.time() - t0
    print(f"{cfg_name} {len(specs)} frames, vis_mode {mode}: {dt * 1e3:.0f} ms, {len(specs) / dt:.1f} frames/s", flush=True)
    be.close()
