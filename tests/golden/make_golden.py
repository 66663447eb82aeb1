"""Generates tests/golden/golden.json with the REFERENCE's own gen_data (oracle/_ref/ref_gen_data = every
reference TU compiled unmodified, see oracle/Makefile).  Only runs where /root/reference exists; the JSON
(sha256 per output file) and a few tiny raw files are committed so that the GPU box -- which has no
reference tree -- can still check against reference output.

    python tests/golden/make_golden.py
"""
import json
import os
import shutil
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle.oracle import Oracle, have_ref  # noqa: E402
from util import cfg_path, dir_digest, make_sequence  # noqa: E402

# name -> (config, scans, beams, azimuth, seed, labels present)
CASES = {
    "small14": ("test_small", 14, 16, 300, 1234, True),
    "small9_seed7": ("test_small", 9, 24, 200, 7, True),
    "example12": ("example", 12, 32, 600, 1234, True),
    "settings6": ("settings", 6, 32, 500, 99, True),
    "kitti8": ("semantic_kitti", 8, 64, 1000, 1234, True),
    "dense5": ("semantic_kitti_dense", 5, 32, 800, 5, True),
    "nolabels5": ("test_small", 5, 16, 200, 3, False),   # label files get created all-zero -> every frame skipped
}


def main():
    if not have_ref():
        raise SystemExit("oracle/_ref is not built: run `make -C oracle ref` where /root/reference exists")
    ref = Oracle("ref")
    out = {}
    for name, (cfg, n, beams, az, seed, labels) in CASES.items():
        tmp = tempfile.mkdtemp(prefix="vxgold_")
        try:
            seq = make_sequence(os.path.join(tmp, "seq"), n, beams, az, seed, labels)
            od = os.path.join(tmp, "out")
            ref.gen_data(cfg_path(cfg), seq, od)
            out[name] = dict(config=cfg, scans=n, beams=beams, azimuth=az, seed=seed, labels=labels, files=dir_digest(od))
            if name == "small14":  # keep one frame raw (a few KB) as a readable fixture
                for ext in ("bin", "label", "occluded", "invalid"):
                    shutil.copy(os.path.join(od, f"000004.{ext}"), os.path.join(HERE, f"small14_000004.{ext}"))
            print(name, len(out[name]["files"]), "files")
        finally:
            shutil.rmtree(tmp, ignore_errors=True)
    with open(os.path.join(HERE, "golden.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
