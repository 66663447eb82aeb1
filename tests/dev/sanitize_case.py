"""Small cases for `compute-sanitizer --tool memcheck|racecheck` (run on a GPU box, one tool per gpurun call):

    compute-sanitizer --tool racecheck python tests/dev/sanitize_case.py > gpurun_out/racecheck.log 2>&1

Both shapes of the traversal kernel, the fill / vote / emit / pack kernels, the public ray, asynchronous submission; every
output is compared with the oracle so that a run the sanitizer slows down 100x still proves it computed the right bytes."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))  # (under tests/: only test code may import the oracle)
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import voxelizer_b200 as vx  # noqa: E402
from frames import assert_frames_equal, oracle_frame, product_frame  # noqa: E402
from oracle.oracle import Oracle  # noqa: E402
from util import cfg_path  # noqa: E402


def main():
    orc = Oracle("port")
    params = vx.synth_params(seed=3, n_beams=16, n_azimuth=200)
    n = 5
    scans, raw = zip(*[vx.synth_scan(params, t) for t in range(n)])
    scans, raw = [s.copy() for s in scans], [l.copy() for l in raw]
    labels = [l & 0xFFFF for l in raw]
    poses = vx.synth_velo_poses(params, n).astype(np.float32)
    for mode in (1, 2):
        be = vx.Backend.from_config(vx.Config(cfg_path("test_small")), debug=True, vis_mode=mode)
        for t in range(n):
            be.upload_scan(t, scans[t], raw[t])
        o = oracle_frame(orc, cfg_path("test_small"), scans, labels, poses, [0], [1, 2, 3, 4])
        p = product_frame(be, poses, [0], [1, 2, 3, 4])
        assert_frames_equal(p, o)
        occ, vis = be.trace_ray(0, be.size[0] - 1, 3, 1, (0.0, 0.0, 0.0))
        assert len(vis) > 0
        print(f"mode {mode}: frame equal to the oracle; public ray visited {len(vis)} cells, occluder {occ}")
        be.close()
    print("SANITIZE CASE OK")


if __name__ == "__main__":
    main()
