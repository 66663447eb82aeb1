"""Developer timing run of the host-buffer path: every step uploads all scans from pinned memory and copies every
output image back to pinned memory.  usage: dev_gpu_e2e.py [n_scans] [reps] [table_sets]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import voxelizer_b200 as vx
from util import cfg_path

n_scans = int(sys.argv[1]) if len(sys.argv) > 1 else 4541
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
table_sets = int(sys.argv[3]) if len(sys.argv) > 3 else 0
CAP = 128000
cfg = vx.Config(cfg_path("semantic_kitti"))
params = vx.synth_params()
px, pl = vx.PinnedBuffer(n_scans * CAP * 16), vx.PinnedBuffer(n_scans * CAP * 4)
counts = vx.synth_generate_many(params, 0, n_scans, px.ptr, pl.ptr, CAP)
poses = vx.synth_velo_poses(params, n_scans).astype(np.float32)
be = vx.Backend.from_config(cfg, table_sets=table_sets)
specs = []
for tgt in vx.plan_targets(cfg, n_scans, poses):
    pb, pe, qb, qe = vx.window(n_scans, cfg.pod.priorScans, cfg.pod.pastScans, int(tgt))
    a = poses[pe - 1]
    ap_p, _ = vx.frame_transforms(a, poses[pb:pe]); ap_q, ep = vx.frame_transforms(a, poses[qb:qe])
    specs.append(vx.FrameSpec(np.arange(pb, pe, dtype=np.uint32), ap_p, np.arange(qb, qe, dtype=np.uint32), ap_q, ep))
per = be.label_bytes + 3 * be.packed_bytes
ring = vx.PinnedBuffer(len(specs) * per)
outs = []
for k in range(len(specs)):
    b = ring.array[k * per:(k + 1) * per]
    outs.append(dict(label=b[:be.label_bytes], bin=b[be.label_bytes:be.label_bytes + be.packed_bytes],
                     occluded=b[be.label_bytes + be.packed_bytes:be.label_bytes + 2 * be.packed_bytes], invalid=b[be.label_bytes + 2 * be.packed_bytes:]))
descs, keep = be.build_descs(specs, outs)

def step():
    t0 = time.time()
    be.evict_all()
    for t in range(n_scans):
        be.upload_scan_ptr(t, px.ptr + t * CAP * 16, pl.ptr + t * CAP * 4, int(counts[t]))
    t1 = time.time()
    be.process_descs(descs, len(specs))
    return t1 - t0, time.time() - t0

step()
be.reset_stage_times()
ts = [step() for _ in range(reps)]
st = be.stage_times()
print(f"table_sets={table_sets}: {len(specs)} frames; enqueue uploads {np.mean([a for a, _ in ts])*1e3:.0f} ms; step {np.mean([b for _, b in ts])*1e3:.0f} ms; "
      f"{len(specs)/np.mean([b for _, b in ts]):.1f} frames/s")
print({k: round(st[k] / reps, 2) for k in ("fill_ms", "vote_ms", "emit_ms", "visibility_ms", "pack_ms", "total_ms")}, "vis launches/step", st["visibility_launches"] / reps)
