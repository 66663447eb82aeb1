#!/usr/bin/env python
"""Developer A/B builds of the CUDA library with other shapes of the throughput traversal kernel
(voxelizer_b200/lib/variants/libvx_<name>.so; select one at run time with VX_B200_CUDA_LIB=<path>).

    python tools/build_variants.py name=threads,maxlive,ctas_per_sm,coarse_words ...
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from voxelizer_b200 import build as vxbuild  # noqa: E402

out_dir = os.path.join(ROOT, "voxelizer_b200", "lib", "variants")
os.makedirs(out_dir, exist_ok=True)
cu = [os.path.join(vxbuild.CSRC, f) for f in ("vx_api.cu", "vx_fill.cu", "vx_visibility.cu")]
procs = []
for spec in sys.argv[1:]:
    name, vals = spec.split("=", 1)
    t, ml, ctas, cw, *rest = vals.split(",")
    extra = [f"-D{x}" for x in rest]  # further macros, e.g. VX_BLOCK=3
    out = os.path.join(out_dir, f"libvx_{name}.so")
    cmd = [vxbuild._nvcc(), *vxbuild.NVCC_FLAGS, f"-DVX_THR_THREADS={t}", f"-DVX_THR_MAXLIVE={ml}", f"-DVX_THR_CTAS={ctas}", f"-DVX_THR_COARSE={cw}", *extra,
           "-shared", "-o", out, *cu]
    procs.append((name, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
for name, p in procs:
    log = p.communicate()[0]
    regs = [ln for ln in log.splitlines() if "registers" in ln or "spill" in ln]
    print(name, "rc", p.returncode, "|", " ".join(x.strip() for x in regs[-4:])[:300])
    if p.returncode:
        print(log[-2000:])
