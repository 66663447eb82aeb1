"""Developer timing run: N frames of semantic_kitti.cfg on a synthetic sequence, stage times and K4 anatomy."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import voxelizer_b200 as vx
from util import cfg_path

n_scans = int(sys.argv[1]) if len(sys.argv) > 1 else 160
cfg_name = sys.argv[2] if len(sys.argv) > 2 else "semantic_kitti"
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
fif = int(sys.argv[4]) if len(sys.argv) > 4 else 0
CAP = 128000
cfg = vx.Config(cfg_path(cfg_name))
params = vx.synth_params()
xyzr = np.empty((n_scans, CAP, 4), np.float32); lab = np.empty((n_scans, CAP), np.uint32)
counts = vx.synth_generate_many(params, 0, n_scans, xyzr.ctypes.data, lab.ctypes.data, CAP)
poses = vx.synth_velo_poses(params, n_scans).astype(np.float32)
be = vx.Backend.from_config(cfg, frames_in_flight=fif, vis_mode=int(os.environ.get("VX_DEV_VIS_MODE", "0")))
for t in range(n_scans):
    be.upload_scan(t, xyzr[t, :counts[t]], lab[t, :counts[t]])
specs = []
for tgt in vx.plan_targets(cfg, n_scans, poses):
    pb, pe, qb, qe = vx.window(n_scans, cfg.pod.priorScans, cfg.pod.pastScans, int(tgt))
    a = poses[pe - 1]
    ap_p, _ = vx.frame_transforms(a, poses[pb:pe]); ap_q, ep = vx.frame_transforms(a, poses[qb:qe])
    specs.append(vx.FrameSpec(np.arange(pb, pe, dtype=np.uint32), ap_p, np.arange(qb, qe, dtype=np.uint32), ap_q, ep))
descs, keep = be.build_descs(specs, None)
be.process_descs(descs, len(specs))
be.reset_stage_times()
t = time.time()
for _ in range(reps):
    be.process_descs(descs, len(specs))
dt = (time.time() - t) / reps
st = be.stage_times()
f = st["frames"]
print(f"{len(specs)} frames/step, {dt*1e3:.1f} ms/step, {len(specs)/dt:.1f} frames/s")
print({k: round(st[k] / reps, 3) for k in ("fill_ms", "vote_ms", "emit_ms", "visibility_ms", "pack_ms", "total_ms")})
print("fill GB/s", 20 * st["points_read"] / (st["fill_ms"] / 1e3) / 1e9, " points/frame", st["points_read"] / f)
cyc = np.array(st["vis_cycles"], float)
print("per frame: rays %.0f steps %.0f waves %.0f resolve %.0f rounds %.0f" % tuple(st[k] / f for k in ("rays_cast", "dda_steps", "vis_waves", "vis_resolve_waves", "vis_rounds")))
print("cycles per wave  A snapshot %.0f | B resolve %.0f | C trace %.0f | D apply %.0f | total %.0f" % (*(cyc[:4] / st["vis_waves"]), cyc[:4].sum() / st["vis_waves"]))
print("Gsteps/s", st["dda_steps"] / (st["visibility_ms"] / 1e3) / 1e9)
if cyc[4] > 0:
    print("trace anatomy (thread 0): longest ray per wave %.1f steps | loop iterations per wave %.1f | attention events per wave %.1f" % tuple(cyc[4:7] / st["vis_waves"]))
