"""Shared helpers of the test-suite."""
from __future__ import annotations

import filecmp
import hashlib
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CONFIGS = os.path.join(ROOT, "configs")
GOLDEN = os.path.join(ROOT, "tests", "golden")


def cfg_path(name: str) -> str:
    return os.path.join(CONFIGS, name + ".cfg")


def make_sequence(directory: str, n_scans: int, beams: int = 16, azimuth: int = 300, seed: int = 1234, labels: bool = True, **kw) -> str:
    import voxelizer_b200 as vx

    p = vx.synth_params(seed=seed, n_beams=beams, n_azimuth=azimuth, **kw)
    vx.write_sequence(p, directory, n_scans, write_labels=labels, threads=4)
    return directory


def compare_dirs(a: str, b: str):
    """List of problems; empty when both directories hold identical files."""
    fa, fb = sorted(os.listdir(a)), sorted(os.listdir(b))
    if fa != fb:
        return [f"file lists differ: {sorted(set(fa) ^ set(fb))[:6]}"]
    return [f for f in fa if not filecmp.cmp(os.path.join(a, f), os.path.join(b, f), shallow=False)]


def dir_digest(d: str) -> dict:
    out = {}
    for f in sorted(os.listdir(d)):
        with open(os.path.join(d, f), "rb") as fh:
            out[f] = hashlib.sha256(fh.read()).hexdigest()
    return out


def load_scan(seq: str, t: int):
    xyzr = np.fromfile(os.path.join(seq, "velodyne", f"{t:06d}.bin"), np.float32).reshape(-1, 4)
    lab = np.fromfile(os.path.join(seq, "labels", f"{t:06d}.label"), np.uint32)
    return xyzr, lab


def random_cfg_text(rng) -> str:
    """A configuration file exercising the parser's corners."""
    lines = [
        "# comment line",
        "   # indented comment",
        f"voxel size: {rng.choice(['0.2', '0.25', ' 0.5 ', '1'])}",
        "min extent: [0, -6.4,-2]",
        "max extent: [ 12.8 , 6.4, 1.2 ]",
        f"prior scans: {rng.integers(0, 3)}",
        f"past scans:{rng.integers(1, 9)}",
        " past scans: 99",            # leading space: key does not match -> "unknown parameter"
        "min range: 2.5",
        "max range: 1e2",
        "bogus key: 7",
        "no colon here",
        "ignore: [0, 2,250 ]",
        "join: [{0: [1]}, {10: [252,13, 16]},{ 20 : [ 256 ] }]",
        "stride num: 2",
        "stride distance: 0.75",
        "",
    ]
    if rng.random() < 0.5:
        lines.append("past distance: 3.9")   # overwrites pastScans (reference bug kept)
    return "\n".join(lines)
