"""Developer timeline of the host-buffer path with several steps in flight: every step uploads all its scans from pinned memory under
its own id range and gets every output image back in its own pinned ring; `depth` steps in flight.  With VX_B200_TRACE=<file> and
VX_B200_FRAME_STATS=<file> the library and this script write a timeline that tools/e2e_timeline.py reads.

    python tools/dev_gpu_e2e_stream.py steps depth:lanes:prio ...      (trace files get the suffix _<depth>_<lanes>_<prio>)
"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
steps = int(sys.argv[1])
settings = sys.argv[2:]
TRACE0, STATS0 = os.environ.get("VX_B200_TRACE"), os.environ.get("VX_B200_FRAME_STATS")
import numpy as np  # noqa: E402

import voxelizer_b200 as vx  # noqa: E402
from util import cfg_path  # noqa: E402

def note(what, a=0, b=0):
    if os.environ.get("VX_B200_TRACE"):
        with open(os.environ["VX_B200_TRACE"], "a") as fp:
            fp.write(f"{what} {time.time_ns()} {a} {b} \n")


n_scans, CAP = 4541, 128000
cfg = vx.Config(cfg_path("semantic_kitti"))
params = vx.synth_params()
px, pl = vx.PinnedBuffer(n_scans * CAP * 16), vx.PinnedBuffer(n_scans * CAP * 4)
counts = vx.synth_generate_many(params, 0, n_scans, px.ptr, pl.ptr, CAP, threads=os.cpu_count() or 8)
poses = vx.synth_velo_poses(params, n_scans).astype(np.float32)
base_specs = []
for tgt in vx.plan_targets(cfg, n_scans, poses):
    pb, pe, qb, qe = vx.window(n_scans, cfg.pod.priorScans, cfg.pod.pastScans, int(tgt))
    a = poses[pe - 1]
    ap_p, _ = vx.frame_transforms(a, poses[pb:pe])
    ap_q, ep = vx.frame_transforms(a, poses[qb:qe])
    base_specs.append((np.arange(pb, pe, dtype=np.uint32), ap_p, np.arange(qb, qe, dtype=np.uint32), ap_q, ep))
for setting in settings:
    depth, lanes, prio = setting.split(":")
    depth = int(depth)
    warmup = max(2, depth)
    os.environ["VX_B200_LANES"] = lanes
    os.environ["VX_B200_FILL_PRIO"] = prio
    if TRACE0:
        os.environ["VX_B200_TRACE"] = f"{TRACE0}_{depth}_{lanes}_{prio}"
    if STATS0:
        os.environ["VX_B200_FRAME_STATS"] = f"{STATS0}_{depth}_{lanes}_{prio}"
    be = vx.Backend.from_config(cfg)
    n = len(base_specs)
    per = be.label_bytes + 3 * be.packed_bytes
    sets, rings = [], []
    for r in range(depth):
        ring = vx.PinnedBuffer(n * per)
        rings.append(ring)
        outs = []
        for k in range(n):
            b = ring.array[k * per:(k + 1) * per]
            outs.append(dict(label=b[:be.label_bytes], bin=b[be.label_bytes:be.label_bytes + be.packed_bytes],
                             occluded=b[be.label_bytes + be.packed_bytes:be.label_bytes + 2 * be.packed_bytes], invalid=b[be.label_bytes + 2 * be.packed_bytes:]))
        off = r << 20
        sets.append(be.build_descs([vx.FrameSpec(p + off, app, q + off, apq, ep) for p, app, q, apq, ep in base_specs], outs))


    def run(k_steps):
        tickets = []
        for k in range(k_steps):
            r = k % depth
            if k >= depth:
                be.wait(tickets[k - depth])
                note("py_wait_done", tickets[k - depth])
            for t in range(n_scans):
                be.evict_scan(t + (r << 20))
            note("py_upload_begin", k)
            for t in range(n_scans):
                be.upload_scan_ptr(t + (r << 20), px.ptr + t * CAP * 16, pl.ptr + t * CAP * 4, int(counts[t]))
            note("py_uploads_enqueued", k)
            tickets.append(be.submit_descs(sets[r][0], n))
        be.wait(0)
        note("py_all_done")


    run(warmup)
    be.sync()
    note("py_timed_begin")
    t0 = time.perf_counter()
    run(steps)
    be.sync()
    dt = (time.perf_counter() - t0) / steps
    print(f"e2e depth {depth} lanes {be.batches_in_flight} prio {prio}: {n / dt:.1f} frames/s ({dt * 1e3:.1f} ms/step over {steps})", flush=True)
    be.close()
    for ring in rings:
        ring.free()
