"""Reads the developer timeline of a streaming run (VX_B200_TRACE + VX_B200_FRAME_STATS files written by the library and by
tools/dev_gpu_e2e_stream.py) and prints, per batch, when it was submitted / enqueued / filled / traced / retired, and how full the
traversal slots of the device were over time.

    python tools/e2e_timeline.py <trace file> <frame stats file> [slots]
"""
import sys

import numpy as np


def main():
    trace, stats = sys.argv[1], sys.argv[2]
    slots = int(sys.argv[3]) if len(sys.argv) > 3 else 1036
    ev = []
    for ln in open(trace):
        p = ln.split()
        if len(p) >= 4:
            ev.append((p[0], int(p[1]), int(p[2]), int(p[3]), " ".join(p[4:])))
    t0 = min(e[1] for e in ev)
    batches = {}
    cur = None
    for ln in open(stats):
        if ln.startswith("# batch"):
            cur = int(ln.split()[2])
            batches[cur] = []
        elif cur is not None:
            f = ln.split()
            batches[cur].append((int(f[12]), int(f[13])))  # start / end, globaltimer ns
    ms = lambda t: (t - t0) / 1e6
    print(f"t = 0 at the first trace event; {len(batches)} batches with frame statistics")
    info = {}
    for what, t, a, b, rest in ev:
        if what in ("enqueue_begin", "enqueue_end", "retire"):
            info.setdefault(a, {})[what] = (t, rest)
    print(" batch | enqueue begin..end ms | first fill -> last emit (device, ms after enqueue end of stage copy) | frames start min..max | "
          "frames end min..max | retire | stage marks")
    for seq in sorted(info):
        d = info[seq]
        fr = np.array(batches.get(seq, []), dtype=np.int64)
        line = f" {seq:5d} |"
        if "enqueue_begin" in d and "enqueue_end" in d:
            line += f" {ms(d['enqueue_begin'][0]):9.1f} .. {ms(d['enqueue_end'][0]):9.1f} |"
        if len(fr):
            line += f" start {ms(fr[:, 0].min()):9.1f} .. {ms(fr[:, 0].max()):9.1f} | end {ms(fr[:, 1].min()):9.1f} .. {ms(fr[:, 1].max()):9.1f} |"
            line += f" mean frame {np.mean(fr[:, 1] - fr[:, 0]) / 1e6:7.1f} ms |"
        if "retire" in d:
            line += f" retire {ms(d['retire'][0]):9.1f} | {d['retire'][1]}"
        print(line)
    print("\n host events")
    for what, t, a, b, rest in ev:
        if what.startswith("py_") or what == "submit":
            print(f"   {ms(t):9.1f} ms  {what} {a}")
    allfr = np.array([x for v in batches.values() for x in v], dtype=np.int64)
    if len(allfr):
        lo, hi = allfr[:, 0].min(), allfr[:, 1].max()
        grid = np.linspace(lo, hi, 61)
        occ = [(int(((allfr[:, 0] <= t) & (allfr[:, 1] > t)).sum())) for t in grid]
        print("\n resident traversal CTAs over time (of", slots, "slots):")
        for t, o in zip(grid, occ):
            print(f"   {ms(int(t)):9.1f} ms  {o:5d}  " + "#" * (o * 60 // slots))
        busy = np.sum(allfr[:, 1] - allfr[:, 0]) / float(hi - lo)
        print(f" mean resident CTAs {busy:.0f} = {busy / slots:.2f} of the slots; {len(allfr)} frames in {(hi - lo) / 1e9:.2f} s = {len(allfr) / ((hi - lo) / 1e9):.1f} frames/s")


if __name__ == "__main__":
    main()
