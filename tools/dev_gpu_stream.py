"""Developer A/B of the batch pipeline: K steps of configs[1] (909 frames each, scans resident) submitted back to back, for a list of
(lanes, fill priority) settings in ONE process (the environment switches are read by vx_create); select another build of the CUDA
library with VX_B200_CUDA_LIB.  Prints one line per setting; with --digest also one sha256 over every output of one step (outputs
into pinned host memory) so that variants can be compared bit for bit.

    python tools/dev_gpu_stream.py [--scans 4541] [--steps 8] [--warmup 4] [--digest] lanes:prio[:leave_room] ...
"""
import argparse
import hashlib
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402

import voxelizer_b200 as vx  # noqa: E402
from util import cfg_path  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--scans", type=int, default=4541)
ap.add_argument("--steps", type=int, default=8)
ap.add_argument("--warmup", type=int, default=4)
ap.add_argument("--config", default="semantic_kitti")
ap.add_argument("--digest", action="store_true")
ap.add_argument("--single", action="store_true", help="also time one awaited step")
ap.add_argument("settings", nargs="*", default=["3:1"])
args = ap.parse_args()

CAP = 128000
cfg = vx.Config(cfg_path(args.config))
params = vx.synth_params()
pin_x = vx.PinnedBuffer(args.scans * CAP * 16)
pin_l = vx.PinnedBuffer(args.scans * CAP * 4)
counts = vx.synth_generate_many(params, 0, args.scans, pin_x.ptr, pin_l.ptr, CAP, threads=os.cpu_count() or 8)
poses = vx.synth_velo_poses(params, args.scans).astype(np.float32)
specs = []
for tgt in vx.plan_targets(cfg, args.scans, poses):
    pb, pe, qb, qe = vx.window(args.scans, cfg.pod.priorScans, cfg.pod.pastScans, int(tgt))
    a = poses[pe - 1]
    ap_p, _ = vx.frame_transforms(a, poses[pb:pe])
    ap_q, ep = vx.frame_transforms(a, poses[qb:qe])
    specs.append(vx.FrameSpec(np.arange(pb, pe, dtype=np.uint32), ap_p, np.arange(qb, qe, dtype=np.uint32), ap_q, ep))
n = len(specs)
print(f"lib {os.environ.get('VX_B200_CUDA_LIB', 'default')}: {args.scans} scans, {n} frames per step", flush=True)

for setting in args.settings:
    parts = setting.split(":")
    os.environ["VX_B200_LANES"] = parts[0]
    os.environ["VX_B200_FILL_PRIO"] = parts[1] if len(parts) > 1 else "1"
    os.environ["VX_B200_LEAVE_ROOM"] = parts[2] if len(parts) > 2 else "0"
    be = vx.Backend.from_config(cfg, vis_mode=int(os.environ.get("VX_DEV_VIS_MODE", "0")))
    lanes = be.batches_in_flight
    for t in range(args.scans):
        be.upload_scan_ptr(t, pin_x.ptr + t * CAP * 16, pin_l.ptr + t * CAP * 4, int(counts[t]))
    descs, keep = be.build_descs(specs, None)

    def run(k):
        tickets = []
        for _ in range(k):
            tickets.append(be.submit_descs(descs, n))
            if len(tickets) > lanes:
                be.wait(tickets[-(lanes + 1)])
        be.wait(0)

    run(args.warmup)
    be.sync()
    be.reset_stage_times()
    t0 = time.perf_counter()
    run(args.steps)
    be.sync()
    dt = (time.perf_counter() - t0) / args.steps
    st = be.stage_times()
    line = (f"lanes {lanes} prio {os.environ['VX_B200_FILL_PRIO']} room {os.environ['VX_B200_LEAVE_ROOM']}: {n / dt:7.1f} frames/s "
            f"({dt * 1e3:7.1f} ms/step over {args.steps}) | per step: fill {st['fill_ms'] / args.steps:6.1f} vote {st['vote_ms'] / args.steps:6.1f} "
            f"emit {st['emit_ms'] / args.steps:6.1f} vis {st['visibility_ms'] / args.steps:7.1f} ms")
    if args.single:
        be.process_descs(descs, n)
        t0 = time.perf_counter()
        be.process_descs(descs, n)
        line += f" | single {n / (time.perf_counter() - t0):6.1f} frames/s"
    if args.digest:
        per = be.label_bytes + 3 * be.packed_bytes
        ring = vx.PinnedBuffer(n * per)
        outs = []
        for i in range(n):
            base = ring.array[i * per:(i + 1) * per]
            outs.append(dict(label=base[: be.label_bytes], bin=base[be.label_bytes: be.label_bytes + be.packed_bytes],
                             occluded=base[be.label_bytes + be.packed_bytes: be.label_bytes + 2 * be.packed_bytes],
                             invalid=base[be.label_bytes + 2 * be.packed_bytes:]))
        d2, keep2 = be.build_descs(specs, outs)
        be.process_descs(d2, n)
        line += " | sha256 " + hashlib.sha256(ring.array[: n * per].tobytes()).hexdigest()[:16]
        ring.free()
    print(line, flush=True)
    be.close()
