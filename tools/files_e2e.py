#!/usr/bin/env python
"""End to end THROUGH FILES ("compute-only and end-to-end (incl. file I/O) frames/s"):
writes a synthetic KITTI-format sequence to a directory, runs the whole gen_data program of this repo on it
(velodyne/*.bin + labels/*.label + poses.txt + calib.txt in, four files per frame out) and prints one JSON line.

    python tools/files_e2e.py [--scans 1000] [--cfg semantic_kitti] [--dir /dev/shm/vx_files_e2e] [--repeat 2]

The reference side of the same measurement is `bench.py --impl reference` (its frames are written as files too).
"""
import argparse
import json
import os
import shutil
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import voxelizer_b200 as vx  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scans", type=int, default=1000)
    ap.add_argument("--cfg", default="semantic_kitti")
    ap.add_argument("--dir", default="/dev/shm/vx_files_e2e")
    ap.add_argument("--repeat", type=int, default=2)
    ap.add_argument("--threads", type=int, default=min(32, os.cpu_count() or 4))
    a = ap.parse_args()

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cfg = os.path.join(root, "configs", a.cfg + ".cfg")
    seq, out = os.path.join(a.dir, "sequence"), os.path.join(a.dir, "out")
    shutil.rmtree(a.dir, ignore_errors=True)
    os.makedirs(seq)
    t0 = time.time()
    vx.write_sequence(vx.synth_params(seed=7), seq, a.scans, write_labels=True, threads=a.threads)
    t_gen = time.time() - t0
    in_bytes = sum(os.path.getsize(os.path.join(dp, f)) for dp, _, fs in os.walk(seq) for f in fs)

    runs = []
    for _ in range(a.repeat):  # the first run also pays CUDA context creation and the first touch of the page cache
        shutil.rmtree(out, ignore_errors=True)
        t0 = time.time()
        st = vx.gen_data(cfg, seq, out)
        wall = time.time() - t0
        out_bytes = sum(os.path.getsize(os.path.join(out, f)) for f in os.listdir(out))
        runs.append(dict(frames=st["frames"], wall_s=round(wall, 3), frames_per_s=round(st["frames"] / wall, 2),
                         pipeline_wall_s=round(st["wall_s"], 3), read_s=round(st["read_s"], 3), gpu_s=round(st["gpu_s"], 3),
                         write_s=round(st["write_s"], 3), out_bytes=out_bytes))
    best = max(runs, key=lambda r: r["frames_per_s"])
    print(json.dumps(dict(metric="frames_per_s_through_files", value=best["frames_per_s"], unit="frames/s", config=a.cfg,
                          scans=a.scans, in_bytes=in_bytes, directory=a.dir, synth_write_s=round(t_gen, 2), runs=runs)))
    shutil.rmtree(a.dir, ignore_errors=True)


if __name__ == "__main__":
    main()
