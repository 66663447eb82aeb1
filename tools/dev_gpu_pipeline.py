"""Developer timing: K steps of the 909-frame job submitted back to back (two batches in flight), stage times."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import voxelizer_b200 as vx
from util import cfg_path

K = int(sys.argv[1]) if len(sys.argv) > 1 else 8
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
descs, keep = be.build_descs(specs, None)
def run(k):
    tickets = []
    for _ in range(k):
        tickets.append(be.submit_descs(descs, len(specs)))
        if len(tickets) > 2:
            be.wait(tickets[-3])
    be.wait(0)
run(2)
be.reset_stage_times()
t = time.time(); run(K); dt = time.time() - t
st = be.stage_times()
print(f"LEAVE_ROOM={os.environ.get('VX_B200_LEAVE_ROOM', '0')}: {K} steps back to back: {dt / K * 1e3:.0f} ms/step, {K * len(specs) / dt:.1f} frames/s; "
      f"per step fill {st['fill_ms'] / K:.0f} vote {st['vote_ms'] / K:.0f} emit {st['emit_ms'] / K:.0f} visibility {st['visibility_ms'] / K:.0f} ms")
