"""In-tree build of the native libraries (no JIT cache: the built .so files travel with the repo
snapshot to the GPU box).

  voxelizer_b200/lib/libvoxelizer_b200.so   CUDA kernels + C ABI (include/voxelizer_b200.h), sm_100a only
  voxelizer_b200/lib/libvxsynth.so          synthetic KITTI sequence generator (CPU)
  voxelizer_b200/lib/libvxhost.so           C++ host layer mirroring the reference API (CPU side)
  voxelizer_b200/bin/vx_gen_data            drop-in for the reference's gen_data executable
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
LIB = os.path.join(PKG, "lib")
BIN = os.path.join(PKG, "bin")
CSRC = os.path.join(PKG, "csrc")
HOST = os.path.join(PKG, "host")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-fmad=false",            # never contract mul+add: the voxel index must match the reference's SSE2 arithmetic
    "-prec-div=true", "-prec-sqrt=true",
    "-Xcompiler", "-fPIC,-ffp-contract=off",
    "-Xptxas", "-v",
]


def _nvcc() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: the CUDA extension cannot be built (there is no CPU fallback)")


def _newer(target: str, sources: list[str]) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def _run(cmd: list[str], log: str | None = None) -> None:
    p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if log:
        with open(log, "w") as f:
            f.write(" ".join(cmd) + "\n" + p.stdout)
    if p.returncode != 0:
        sys.stderr.write(p.stdout)
        raise RuntimeError("build failed: " + " ".join(cmd))


def build_cuda(force: bool = False, out: str | None = None) -> str:
    os.makedirs(LIB, exist_ok=True)
    out = out or os.path.join(LIB, "libvoxelizer_b200.so")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    cu = [os.path.join(CSRC, f) for f in ("vx_api.cu", "vx_fill.cu", "vx_visibility.cu")]
    deps = cu + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh"))]
    deps.append(os.path.join(ROOT, "include", "voxelizer_b200.h"))
    if force or _newer(out, deps):
        _run([_nvcc(), *NVCC_FLAGS, "-shared", "-o", out, *cu], log=os.path.join(LIB, "nvcc_build.log"))
    return out


def build_synth(force: bool = False) -> str:
    os.makedirs(LIB, exist_ok=True)
    out = os.path.join(LIB, "libvxsynth.so")
    src = os.path.join(PKG, "synth", "synth.cpp")
    if force or _newer(out, [src]):
        _run(["g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-pthread", src, "-o", out])
    return out


def build_host(force: bool = False) -> list[str]:
    outs = []
    if not os.path.isdir(HOST):
        return outs
    srcs = sorted(os.path.join(HOST, f) for f in os.listdir(HOST) if f.endswith(".cpp") and f != "gen_data.cpp")
    hdrs = [os.path.join(HOST, f) for f in os.listdir(HOST) if f.endswith(".h")]
    if not srcs:
        return outs
    os.makedirs(LIB, exist_ok=True)
    os.makedirs(BIN, exist_ok=True)
    lib = os.path.join(LIB, "libvxhost.so")
    cuda_lib = os.path.join(LIB, "libvoxelizer_b200.so")
    common = ["g++", "-std=c++17", "-O2", "-fPIC", "-ffp-contract=off", "-pthread", "-I" + os.path.join(ROOT, "include"), "-I" + HOST]
    if force or _newer(lib, srcs + hdrs + [cuda_lib]):
        _run([*common, "-shared", *srcs, "-o", lib, "-L" + LIB, "-lvoxelizer_b200", "-Wl,-rpath,$ORIGIN"])
    outs.append(lib)
    main = os.path.join(HOST, "gen_data.cpp")
    if os.path.exists(main):
        exe = os.path.join(BIN, "vx_gen_data")
        if force or _newer(exe, [main, lib] + hdrs):
            _run([*common, main, "-o", exe, "-L" + LIB, "-lvxhost", "-lvoxelizer_b200", "-Wl,-rpath,$ORIGIN/../lib"])
        outs.append(exe)
    return outs


def build_all(force: bool = False) -> None:
    build_cuda(force)
    build_synth(force)
    build_host(force)


if __name__ == "__main__":
    build_all(force="--force" in sys.argv)
    print("built:", sorted(os.listdir(LIB)))
