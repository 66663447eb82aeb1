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


# ------------------------------------------------------------------------------------------------------
# tests/callers/libvxcallers.so: callers shaped like the reference's gen_data / Mainframe, written against the kept host API
# ------------------------------------------------------------------------------------------------------
CALLERS_SRC = os.path.join(ROOT, "tests", "callers", "callers.cpp")
CALLERS_LIB = os.path.join(ROOT, "tests", "callers", "libvxcallers.so")
_callers = None


def build_callers(force: bool = False) -> str:
    import subprocess
    host = os.path.join(ROOT, "voxelizer_b200", "host")
    lib = os.path.join(ROOT, "voxelizer_b200", "lib")
    deps = [CALLERS_SRC, os.path.join(lib, "libvxhost.so")] + [os.path.join(host, f) for f in os.listdir(host) if f.endswith(".h")]
    if force or not os.path.exists(CALLERS_LIB) or os.path.getmtime(CALLERS_LIB) < max(os.path.getmtime(d) for d in deps):
        subprocess.check_call(["g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-ffp-contract=off", "-pthread", "-I" + os.path.join(ROOT, "include"),
                               "-I" + host, CALLERS_SRC, "-o", CALLERS_LIB, "-L" + lib, "-lvxhost", "-lvoxelizer_b200",
                               "-Wl,-rpath,$ORIGIN/../../voxelizer_b200/lib"])
    return CALLERS_LIB


def callers_lib():
    global _callers
    if _callers is None:
        import ctypes as C
        import voxelizer_b200 as vx
        vx.host_lib()
        L = C.CDLL(build_callers())
        L.vxt_last_error.restype = C.c_char_p
        L.vxt_gen_data_frame_by_frame.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, C.c_int, C.c_int64, C.POINTER(C.c_uint64)]
        L.vxt_build_voxel_grids.argtypes = [C.c_char_p, C.c_char_p, C.c_int32, C.c_int] + [C.c_void_p] * 5 + [C.c_uint64, C.c_void_p, C.POINTER(C.c_double)]
        L.vxt_voxel_grid_probe.argtypes = ([C.c_int, C.c_float] + [C.c_void_p] * 4 + [C.c_uint32, C.c_void_p, C.c_uint32] + [C.c_void_p] * 8 +
                                           [C.c_uint32] + [C.c_void_p] * 4 + [C.c_uint64, C.c_void_p, C.c_int] + [C.c_void_p] * 4)
        _callers = L
    return _callers


def _ptr(a):
    import ctypes as C
    return a.ctypes.data_as(C.c_void_p)


def gen_data_frame_by_frame(cfg: str, seq_dir: str, out_dir: str, device: int = 0, max_frames: int = -1) -> int:
    """The batch program written against VoxelGrid / fillVoxelGrid / saveVoxelGrid, one frame at a time."""
    import ctypes as C
    L = callers_lib()
    os.makedirs(out_dir, exist_ok=True)
    n = C.c_uint64(0)
    if L.vxt_gen_data_frame_by_frame(cfg.encode(), seq_dir.encode(), out_dir.encode(), device, max_frames, C.byref(n)) != 0:
        raise RuntimeError(L.vxt_last_error().decode())
    return int(n.value)


def build_voxel_grids(cfg: str, seq_dir: str, target: int, nvox: int, device: int = 0) -> dict:
    """Mainframe::buildVoxelGrids-shaped caller: both grids of one target + the voxel lists the GUI draws."""
    import ctypes as C
    L = callers_lib()
    lp, lq = np.zeros((nvox, 4), np.uint32), np.zeros((nvox, 4), np.uint32)
    op, oq, iq = (np.zeros(nvox, np.uint32) for _ in range(3))
    counts = np.zeros(5, np.uint64)
    sec = C.c_double()
    if L.vxt_build_voxel_grids(cfg.encode(), seq_dir.encode(), target, device, _ptr(lp), _ptr(lq), _ptr(op), _ptr(oq), _ptr(iq), nvox,
                               _ptr(counts), C.byref(sec)) != 0:
        raise RuntimeError(L.vxt_last_error().decode())
    c = [int(x) for x in counts]
    return dict(labeled_prior=lp[: c[0]], labeled_past=lq[: c[1]], occluded_prior=op[: c[2]], occluded_past=oq[: c[3]], invalid_past=iq[: c[4]],
                seconds=sec.value)


def voxel_grid_probe(resolution, min_extent, max_extent, points4, labels, endpoints, nvox: int, rays=(), device: int = 0,
                     insert_occlusion_labels: bool = False) -> dict:
    """vxb::VoxelGrid used directly (raw insert, updateOcclusions, updateInvalid, every per-voxel query, public
    occludedBy rays).  Per-voxel arrays come back in the reference's internal index order i + j*Sx + k*Sx*Sy.
    rays: list of ((i, j, k), (ex, ey, ez))."""
    L = callers_lib()
    f32 = lambda a: np.ascontiguousarray(a, np.float32)
    pts = f32(points4).reshape(-1, 4)
    lab = np.ascontiguousarray(labels, np.uint32)
    ep = f32(endpoints).reshape(-1, 3)
    mn, mx = f32(list(min_extent)[:3]), f32(list(max_extent)[:3])
    size, off = np.zeros(3, np.uint32), np.zeros(3, np.float32)
    out = {k: np.zeros(nvox, np.uint8) for k in ("occluded", "free", "invalid")}
    out.update({k: np.zeros(nvox, np.uint32) for k in ("count", "n_labels", "majority")})
    images = dict(label=np.zeros(nvox, np.uint16), bin=np.zeros(nvox // 8, np.uint8), occluded=np.zeros(nvox // 8, np.uint8),
                  invalid=np.zeros(nvox // 8, np.uint8))
    n_rays = len(rays)
    rs = np.ascontiguousarray([r[0] for r in rays], np.int32).reshape(-1, 3)
    re_ = f32([r[1] for r in rays]).reshape(-1, 3)
    occluder = np.zeros(max(1, n_rays), np.int32)
    cap = 4096 * max(1, n_rays)
    visited = np.zeros((cap, 3), np.int32)
    offsets = np.zeros(n_rays + 1, np.uint64)
    rc = L.vxt_voxel_grid_probe(device, resolution, _ptr(mn), _ptr(mx), _ptr(pts), _ptr(lab), len(lab), _ptr(ep), len(ep), _ptr(size), _ptr(off),
                                _ptr(out["occluded"]), _ptr(out["free"]), _ptr(out["invalid"]), _ptr(out["count"]), _ptr(out["n_labels"]),
                                _ptr(out["majority"]), n_rays, _ptr(rs), _ptr(re_), _ptr(occluder), _ptr(visited), cap, _ptr(offsets),
                                int(insert_occlusion_labels), _ptr(images["label"]), _ptr(images["bin"]), _ptr(images["occluded"]),
                                _ptr(images["invalid"]))
    if rc != 0:
        raise RuntimeError(L.vxt_last_error().decode())
    out["size"] = tuple(int(x) for x in size)
    out["offset"] = off
    out["images"] = images
    out["rays"] = [(int(occluder[r]), visited[int(offsets[r]): int(offsets[r + 1])].copy()) for r in range(n_rays)]
    return out
