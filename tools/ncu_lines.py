#!/usr/bin/env python
"""Per-source-line instruction counts and stall samples of one kernel of an .ncu-rep (needs -lineinfo
and `--import-source on`).  ncu's CSV source page is SASS-only, so the SASS offsets are joined with
nvdisasm's line table of the cubin inside the built library.

    python tools/ncu_lines.py gpurun_out/X.ncu-rep vx_visibility [top]
"""
import csv
import io
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "voxelizer_b200", "lib", "libvoxelizer_b200.so")


def line_table(kernel):
    tmp = tempfile.mkdtemp(prefix="vxsass_")
    subprocess.run(["cuobjdump", "-xelf", "all", LIB], cwd=tmp, capture_output=True)
    table = {}
    for f in sorted(os.listdir(tmp)):
        if not f.endswith(".cubin") or f.count("-") > 0:
            continue
        out = subprocess.run(["nvdisasm", "-g", os.path.join(tmp, f)], capture_output=True, text=True).stdout
        if kernel not in out:
            continue
        cur, infn = None, False
        for ln in out.split("\n"):
            m = re.match(r"\s*\.text\.(\S+):", ln)
            if m:
                infn = kernel in m.group(1)
                continue
            if not infn:
                continue
            m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
            if m:
                cur = (os.path.basename(m.group(1)), int(m.group(2)))
                continue
            m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+\S", ln)
            if m and cur:
                table[int(m.group(1), 16)] = cur
    return table


def main():
    rep, kernel = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kernel}"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    idx = {n: i for i, n in enumerate(rows[hi])}
    table = line_table(kernel)
    base = int(rows[hi + 1][0], 16)
    agg, tot, tots = {}, 0, 0
    stall_cols = [c for c in idx if c.startswith("stall_") and "Not Issued" not in c]
    for r in rows[hi + 1:]:
        if len(r) <= idx["# Samples"] or not r[0].startswith("0x"):
            continue
        off = int(r[0], 16) - base
        ie, smp = int(r[idx["Instructions Executed"]]), int(r[idx["# Samples"]])
        key = table.get(off, ("?", 0))
        a = agg.setdefault(key, [0, 0, {}])
        a[0] += ie
        a[1] += smp
        for c in stall_cols:
            v = int(r[idx[c]] or 0)
            if v:
                a[2][c] = a[2].get(c, 0) + v
        tot += ie
        tots += smp
    src = {}
    for f in os.listdir(os.path.join(ROOT, "voxelizer_b200", "csrc")):
        src[f] = open(os.path.join(ROOT, "voxelizer_b200", "csrc", f)).read().split("\n")
    print(f"kernel {kernel}: {tot} warp instructions, {tots} stall samples")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        line = src[k[0]][k[1] - 1].strip()[:80] if k[0] in src and 0 < k[1] <= len(src[k[0]]) else ""
        st = ",".join(f"{c[6:]}:{n * 100 // max(1, v[1])}" for c, n in sorted(v[2].items(), key=lambda x: -x[1])[:3])
        print(f"{v[1] / max(1, tots) * 100:5.1f}% smp {v[0] / max(1, tot) * 100:5.1f}% inst  {k[0]}:{k[1]:<4} [{st}]  {line}")


if __name__ == "__main__":
    main()
