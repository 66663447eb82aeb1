#!/usr/bin/env python
"""DRAM bytes of ONE vx_fill_kernel launch (an `ncu --set full` capture) next to the algorithmic bytes of that launch.

    python tools/fill_traffic.py gpurun_out/fill.ncu-rep gpurun_out/plain.json <launch index> profiles/rNN_fill_traffic.json

plain.json is the bench line of the SAME command run without ncu: its roofline.bytes_of_launch lists the algorithmic
bytes (20 B x window points) of every fill launch of one step in launch order; <launch index> says which of them the
capture holds (ncu -k regex:vx_fill -s <index> -c 1).  bench.py reads the newest profiles/*_fill_traffic.json for
roofline.traffic instead of a literal."""
import csv
import io
import json
import subprocess
import sys


def main():
    rep, plain, index, out = sys.argv[1], sys.argv[2], int(sys.argv[3]), sys.argv[4]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    head, units = rows[0], rows[1]
    d = dict(zip(head, rows[2]))
    u = dict(zip(head, units))

    def to_bytes(key):
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}[u[key]]
        return float(d[key]) * scale

    dram = to_bytes("dram__bytes_read.sum") + to_bytes("dram__bytes_write.sum")
    line = json.loads(open(plain).read().strip().splitlines()[-1])
    alg = line["roofline"]["bytes_of_launch"][index]
    rec = {"kernel": d.get("Kernel Name", "vx_fill_kernel"), "launch": f"fill launch {index} of one step of `{' '.join(sys.argv[5:]) or 'bench.py'}`",
           "grid": d.get("launch__grid_size"), "duration": f"{d.get('gpu__time_duration.sum')} {u.get('gpu__time_duration.sum')} (under ncu)",
           "dram_bytes": dram, "algorithmic_bytes": alg, "ratio": dram / alg, "source_report": rep}
    json.dump(rec, open(out, "w"), indent=1)
    print(json.dumps(rec, indent=1))


if __name__ == "__main__":
    main()
