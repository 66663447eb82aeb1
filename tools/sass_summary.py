#!/usr/bin/env python
"""Tracked static summary of the built CUDA library (no SASS listing was tracked): per kernel the ptxas
resource line (registers, spills, shared memory; from voxelizer_b200/lib/nvcc_build.log) and a histogram of SASS mnemonics
(cuobjdump -sass), with the mnemonics that matter for the B200 checks called out (REDG / ATOMG, UBLKCP / UTMALDG, LDGSTS, BAR).

    python tools/sass_summary.py > profiles/r02_sass_summary.txt
"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "voxelizer_b200", "lib", "libvoxelizer_b200.so")
LOG = os.path.join(ROOT, "voxelizer_b200", "lib", "nvcc_build.log")


def demangle(name):
    m = re.search(r"\d+(vx_[a-z_]+_kernel)", name)
    cfg = re.search(r"VisCfgI((?:Li\d+E)+)", name)
    shape = "<" + ",".join(re.findall(r"Li(\d+)E", cfg.group(1))) + ">" if cfg else ""
    return (m.group(1) if m else name) + shape


def main():
    print("# ptxas -v (voxelizer_b200/lib/nvcc_build.log)")
    log = open(LOG).read().splitlines()
    print(log[0].replace(ROOT + "/", ""))
    cur = None
    for ln in log:
        m = re.search(r"Compiling entry function '([^']+)'", ln)
        if m:
            cur = demangle(m.group(1))
        elif cur and ("Used" in ln or "spill" in ln):
            print(f"{cur:60s} {ln.replace('ptxas info    :', '').strip()}")
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    print("\n# SASS (cuobjdump -sass): instructions per kernel, most frequent mnemonics, and the ones the B200 checks look for")
    arch = re.search(r"arch = (sm_\w+)", sass)
    print("arch:", arch.group(1) if arch else "?")
    kernels = collections.OrderedDict()
    cur = None
    for ln in sass.splitlines():
        m = re.search(r"Function : (\S+)", ln)
        if m:
            cur = demangle(m.group(1))
            kernels[cur] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", ln)
        if m and cur:
            kernels[cur][m.group(1)] += 1
    watch = ("REDG", "RED", "ATOMG", "ATOM", "ATOMS", "UBLKCP", "UTMALDG", "LDGSTS", "BAR", "LDL", "STL", "LDG", "STG", "LDS", "STS", "SHFL", "VOTE")
    for k, c in kernels.items():
        total = sum(c.values())
        top = ", ".join(f"{m} {n}" for m, n in c.most_common(8))
        seen = ", ".join(f"{m} {c[m]}" for m in watch if c[m])
        print(f"{k}: {total} instructions | top: {top} | {seen}")


if __name__ == "__main__":
    main()
