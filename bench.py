#!/usr/bin/env python
"""Benchmark of the gen_data hot path (BASELINE.json: voxelized frames/s, 256x256x32 @ 0.2 m).

    python bench.py [--gpus N] [--steps K] [--warmup W]            # this repo's CUDA path
    python bench.py --impl reference [--steps K] [--warmup W]      # the reference's CPU path on host cores

A STEP is one pass of the hot path over one whole synthetic sequence: 4541 HDL-64-like scans
(SemanticKITTI seq-00 shape), configs/semantic_kitti.cfg => 909 target frames, each aggregating its
window of up to 71 scans (BASELINE.json configs[1]).  With N GPUs every rank runs its own sequence
(seed 1234 + rank): frames shard naturally, there is no data-path collective (weak scaling); the only
communication is the barrier and the max-reduction of the per-rank device times.

value : frames/s with all scans already resident in HBM and outputs left in HBM.
e2e   : frames/s through the C ABI with HOST buffers: every step uploads every scan from pinned host
        memory and copies every output file image back to pinned host memory.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "voxelized frames/sec (256x256x32 @0.2m)"
UNIT = "frames/s"
CAP_POINTS = 128000  # 64 beams x 2000 azimuth steps


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--scans", type=int, default=4541)
    ap.add_argument("--config", default="semantic_kitti")
    ap.add_argument("--frames-in-flight", type=int, default=0)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--ref-cores", type=int, default=0)
    return ap.parse_args()


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------------------
# reference arm: the reference's own CPU code (oracle/_ref when it was compiled, else the port)
# ------------------------------------------------------------------------------------------------------
_REF_DATA = {}  # inherited by the forked workers (no pickling of the scans)


def _ref_worker(target):
    kind, cfg_file, scans, labels, poses, n_past = (_REF_DATA[k] for k in ("kind", "cfg", "scans", "labels", "poses", "n_past"))
    from oracle.oracle import Oracle
    import tempfile
    orc = Oracle(kind)
    cfg = orc.load_config(cfg_file)
    p = cfg.pod
    ids = list(range(target, min(len(scans), target + n_past + 1)))
    t0 = time.perf_counter()
    gi = orc.grid(p.voxelSize, list(p.minExtent), list(p.maxExtent))
    gt = orc.grid(p.voxelSize, list(p.minExtent), list(p.maxExtent))
    anchor = poses[target]
    pick = lambda idx: ([scans[i] for i in idx], [labels[i] for i in idx], np.stack([poses[i] for i in idx]))
    ps, pl, pp = pick(ids[:1])
    qs, ql, qp = pick(ids[1:])
    gi.fill(cfg, anchor, ps, pl, pp)                      # gen_data.cpp:96-98
    gt.fill(cfg, anchor, ps, pl, pp)
    gt.fill(cfg, anchor, qs, ql, qp)
    gi.update_occlusions()                                # gen_data.cpp:102-103
    gt.update_occlusions()
    for i in ids[1:]:                                     # gen_data.cpp:113-116
        gt.update_invalid(orc.endpoint(anchor, poses[i]))
    with tempfile.TemporaryDirectory(dir="/dev/shm" if os.path.isdir("/dev/shm") else None) as d:
        gi.save(d, "%06d" % target, "input")              # gen_data.cpp:120-121
        gt.save(d, "%06d" % target, "target")
    return time.perf_counter() - t0


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import multiprocessing as mp

    import voxelizer_b200 as vx
    from oracle.oracle import have_ref
    from util import cfg_path

    kind = "ref" if have_ref() else "port"
    cores = args.ref_cores or (os.cpu_count() or 1)
    cfg = vx.Config(cfg_path(args.config))
    stride = max(1, cfg.pod.stride_num)
    n_past = cfg.pod.pastScans
    n_frames = cores
    n_scans = stride * (n_frames - 1) + n_past + 1
    params = vx.synth_params(seed=1234)
    xyzr = np.empty((n_scans, CAP_POINTS, 4), np.float32)
    lab = np.empty((n_scans, CAP_POINTS), np.uint32)
    counts = vx.synth_generate_many(params, 0, n_scans, xyzr.ctypes.data, lab.ctypes.data, CAP_POINTS)
    scans = [xyzr[t, : counts[t]] for t in range(n_scans)]
    labels = [lab[t, : counts[t]] & 0xFFFF for t in range(n_scans)]
    poses = vx.synth_velo_poses(params, n_scans).astype(np.float32)
    _REF_DATA.update(kind=kind, cfg=cfg_path(args.config), scans=scans, labels=labels, poses=poses, n_past=n_past)
    jobs = [stride * i for i in range(n_frames)]
    step_s = []
    ctx = mp.get_context("fork")
    with ctx.Pool(cores) as pool:
        for s in range(args.warmup + args.steps):
            t0 = time.perf_counter()
            per_frame = pool.map(_ref_worker, jobs, chunksize=1)
            dt = time.perf_counter() - t0
            if s >= args.warmup:
                step_s.append(dt)
    ms = 1e3 * float(np.mean(step_s))
    value = n_frames / (ms / 1e3)
    sample = (f"{n_frames} full-window frames ({n_past + 1} scans each, targets 0,{stride},..) of the same synthetic sequence per step, "
              f"one frame per process on {cores} host processes; fill + updateOcclusions x2 + updateInvalid x{n_past} + saveVoxelGrid "
              f"to tmpfs; mean single-frame time {np.mean(per_frame):.2f} s")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32+u32", "data": "synthetic",
        "config": {"workload": f"{args.config}.cfg, synthetic seq-00-shaped sequence, full {n_past + 1}-scan windows", "frames_per_step": n_frames},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "reference" if kind == "ref" else "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    def __init__(self, device_index):
        super().__init__(daemon=True)
        self.idx = device_index
        self.samples = []
        self.stop_flag = threading.Event()

    def run(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.idx}", f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            self.stop_flag.wait(0.2)

    def summary(self):
        sm = [float(s[0]) for s in self.samples if s and s[0].replace(".", "").isdigit()]
        mx = [float(s[1]) for s in self.samples if len(s) > 1 and s[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for s in self.samples for i in range(4) if len(s) > 2 + i and s[2 + i].lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(self.samples)}


def run_ours(args):
    import torch
    import torch.distributed as dist

    import voxelizer_b200 as vx
    from util import cfg_path

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: voxelizer_b200 has no CPU path")
    torch.cuda.set_device(local_rank)
    if world > 1:
        # stdout carries exactly one JSON line: NCCL prints its version banner there while the communicator is
        # created (first collective), so fd 1 points at stderr until that has happened
        os.environ.setdefault("NCCL_DEBUG", "WARN")
        sys.stdout.flush()
        saved_fd = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_fd, 1)
            os.close(saved_fd)

    # ---- CPU baseline (rank 0, N=1 only): the reference arm as a subprocess, one step
    cpu_baseline = None
    if world == 1 and not args.no_cpu_baseline:
        try:
            out = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "reference", "--steps", "1", "--warmup", "0",
                                  "--config", args.config], capture_output=True, text=True, timeout=900)
            cpu_baseline = json.loads(out.stdout.strip().splitlines()[-1])["cpu_baseline"]
        except Exception as e:  # the baseline is reported, never required for the GPU number
            cpu_baseline = {"value": None, "unit": UNIT, "cores": 0, "kind": "port", "sample": f"failed: {e}"}

    # ---- synthetic sequence of this rank, straight into pinned host memory
    cfg = vx.Config(cfg_path(args.config))
    n_scans = args.scans
    params = vx.synth_params(seed=1234 + rank)
    t_gen = time.perf_counter()
    pin_xyzr = vx.PinnedBuffer(n_scans * CAP_POINTS * 16)
    pin_lab = vx.PinnedBuffer(n_scans * CAP_POINTS * 4)
    counts = vx.synth_generate_many(params, 0, n_scans, pin_xyzr.ptr, pin_lab.ptr, CAP_POINTS, threads=max(1, (os.cpu_count() or 1) // max(1, world)))
    poses = vx.synth_velo_poses(params, n_scans).astype(np.float32)
    lab_view = pin_lab.array.view(np.uint32).reshape(n_scans, CAP_POINTS)
    labelled = np.array([int(((lab_view[t, : counts[t]] & 0xFFFF) > 0).sum()) for t in range(n_scans)], np.uint32)
    t_gen = time.perf_counter() - t_gen

    be = vx.Backend.from_config(cfg, device=local_rank, frames_in_flight=args.frames_in_flight,
                                stream=torch.cuda.current_stream().cuda_stream)

    def upload_all():
        for t in range(n_scans):
            be.upload_scan_ptr(t, pin_xyzr.ptr + t * CAP_POINTS * 16, pin_lab.ptr + t * CAP_POINTS * 4, int(counts[t]))

    # ---- plan exactly like gen_data: targets, windows, gate, transforms
    targets = vx.plan_targets(cfg, n_scans, poses)
    specs, window_points = [], 0
    for tgt in targets:
        pb, pe, qb, qe = vx.window(n_scans, cfg.pod.priorScans, cfg.pod.pastScans, int(tgt))
        pts = np.uint32(counts[qb:qe].sum(dtype=np.uint32))
        labc = np.uint32(labelled[qb:qe].sum(dtype=np.uint32))
        if not (np.float32(100.0) * np.float32(labc) / np.float32(pts) > 0):
            continue  # "skipped."
        anchor = poses[pe - 1]
        ap_p, _ = vx.frame_transforms(anchor, poses[pb:pe])
        ap_q, ep = vx.frame_transforms(anchor, poses[qb:qe])
        specs.append(vx.FrameSpec(np.arange(pb, pe, dtype=np.uint32), ap_p, np.arange(qb, qe, dtype=np.uint32), ap_q, ep))
        window_points += int(counts[pb:pe].sum()) + int(counts[qb:qe].sum())
    n_frames = len(specs)
    descs_dev, keep1 = be.build_descs(specs, None)  # outputs stay in HBM

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        w0 = time.perf_counter()
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        wall = time.perf_counter() - w0
        dev_ms = e0.elapsed_time(e1)
        t = torch.tensor([dev_ms, wall * 1e3], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        barrier()
        return float(t[0]), float(t[1])

    # ---- value: inputs resident in HBM
    upload_all()
    step_dev = lambda: be.process_descs(descs_dev, n_frames)
    for _ in range(args.warmup):
        step_dev()
    be.reset_stage_times()
    sampler = ClockSampler(local_rank)
    sampler.start()
    dev_ms, wall_ms = timed(step_dev, args.steps)
    sampler.stop_flag.set()
    sampler.join(timeout=5)
    st = be.stage_times()
    ms_per_step = dev_ms / args.steps
    total_frames = n_frames * world
    if world > 1:
        tf = torch.tensor([n_frames], dtype=torch.int64, device="cuda")
        dist.all_reduce(tf)
        total_frames = int(tf)
    value = total_frames / (ms_per_step / 1e3)

    # ---- e2e: host buffers, H2D of every scan and D2H of every output inside the timed region
    e2e = None
    ring = None
    if not args.no_e2e:
        try:
            ring = vx.PinnedBuffer(n_frames * (be.label_bytes + 3 * be.packed_bytes))
        except Exception as exc:  # the number is optional, the run is not
            sys.stderr.write(f"bench: no pinned output ring ({exc}); e2e skipped\n")
        if world > 1:  # the timed loop is collective: every rank measures e2e, or none does
            ok = torch.tensor([1 if ring is not None else 0], dtype=torch.int32, device="cuda")
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
            if int(ok) == 0:
                ring = None
    if ring is not None:
        per = be.label_bytes + 3 * be.packed_bytes
        outs = []
        for k in range(n_frames):
            base = ring.array[k * per:(k + 1) * per]
            outs.append(dict(label=base[: be.label_bytes], bin=base[be.label_bytes: be.label_bytes + be.packed_bytes],
                             occluded=base[be.label_bytes + be.packed_bytes: be.label_bytes + 2 * be.packed_bytes],
                             invalid=base[be.label_bytes + 2 * be.packed_bytes:]))
        descs_host, keep2 = be.build_descs(specs, outs)

        def step_e2e():
            be.evict_all()
            upload_all()
            be.process_descs(descs_host, n_frames)

        step_e2e()
        e_dev_ms, e_wall_ms = timed(step_e2e, max(1, args.steps))
        e_ms = e_wall_ms / max(1, args.steps)  # includes host-side staging gaps
        h2d = int(counts.sum()) * 20 + sum((len(s.prior_ids) + len(s.past_ids)) * (64 + 24) + len(s.endpoints) * 12 for s in specs)
        d2h = n_frames * per
        e2e = {"value": total_frames / (e_ms / 1e3), "unit": UNIT, "h2d_bytes_per_step": h2d * world, "d2h_bytes_per_step": d2h * world,
               "ms_per_step": e_ms, "device_ms_per_step": e_dev_ms / max(1, args.steps)}

    if rank == 0:
        peak, peak_src = measured_peak_gbs()
        nvox = be.nvox
        launches = {k: int(st[k]) for k in ("fill_launches", "vote_launches", "emit_launches", "visibility_launches", "pack_launches")}
        fill_bytes = 20.0 * st["points_read"]                                   # xyzr 16 B + label 4 B per window point
        fill_s = st["fill_ms"] / 1e3
        fill_gbs = fill_bytes / fill_s / 1e9 if fill_s > 0 else 0.0
        ivp_bytes = fill_bytes + st["frames"] * (2.0 * nvox + 3.0 * nvox / 8.0)  # SURVEY.md 8(d) B_frame
        ivp_s = (st["fill_ms"] + st["vote_ms"] + st["emit_ms"] + st["pack_ms"]) / 1e3
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32+u32",
            "data": "synthetic",
            "config": {"workload": f"{args.config}.cfg on a synthetic {n_scans}-scan sequence per GPU (seq-00 shape, ~{int(counts.mean())} pts/scan): "
                                   f"{n_frames} target frames per step per GPU, window up to {cfg.pod.pastScans + 1} scans",
                       "frames_per_step_per_gpu": n_frames, "frames_in_flight": be.frames_in_flight,
                       "l2_policy": "inputs larger than L2 (window of one frame = 180 MB, sequence = 11.6 GB)",
                       "parallelism": f"{world} x independent target-range shards, no data-path collective"},
            "wall_ms_per_step": wall_ms / args.steps,
            "gpu_launches": sum(launches.values()),
            "launches": launches,
            "stage_ms_per_step": {k: st[k] / args.steps for k in ("fill_ms", "vote_ms", "emit_ms", "visibility_ms", "pack_ms", "total_ms")},
            "roofline": {"bound": "hbm", "kernel": "vx_fill_kernel (K1+K2: transform, filter, key, histogram insert)",
                         "achieved": fill_gbs, "peak": peak, "unit": "GB/s", "frac": fill_gbs / peak,
                         # dram bytes per launch: the ncu --set full capture in profiles/ measured 0.142 x the algorithmic
                         # bytes (a scan tile is loaded once and inserted into every frame of the launch that uses it)
                         "traffic": 0.142 * fill_bytes / max(1, launches["fill_launches"]),
                         "traffic_source": "profiles/r01_ncu_full_fill_final.txt (dram__bytes_read+write / algorithmic bytes of that launch)",
                         "peak_source": peak_src, "bytes_per_launch": fill_bytes / max(1, launches["fill_launches"]),
                         "ms_per_launch": st["fill_ms"] / max(1, launches["fill_launches"]),
                         "insert_vote_pack": {"achieved": ivp_bytes / ivp_s / 1e9 if ivp_s > 0 else 0.0,
                                              "frac": (ivp_bytes / ivp_s / 1e9 / peak) if ivp_s > 0 else 0.0,
                                              "bytes_per_frame": ivp_bytes / max(1, st["frames"])}},
            "traversal": {"kernel": "vx_visibility_kernel (K4)", "rays_per_frame": st["rays_cast"] / max(1, st["frames"]),
                          "dda_steps_per_frame": st["dda_steps"] / max(1, st["frames"]),
                          "gsteps_per_s": st["dda_steps"] / (st["visibility_ms"] / 1e3) / 1e9 if st["visibility_ms"] > 0 else 0.0},
            "cpu_baseline": cpu_baseline,
            "e2e": e2e,
            "clocks": sampler.summary(),
            "setup_s": {"generate": t_gen},
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    args = parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
