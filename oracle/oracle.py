"""ORACLE / TEST INFRASTRUCTURE ONLY.

ctypes front-end over the two CPU checkers:

* ``port``  -- oracle/liboracle_port.so, the from-scratch restatement (oracle/port.cpp);
* ``ref``   -- oracle/_ref/libvx_ref.so, the reference's own translation units compiled unmodified
               (oracle/Makefile, oracle/ref_bridge.cpp).  Exists only where it was built from
               /root/reference (this container) or travelled with the gpurun snapshot.

Only tests/, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference`` legs may
import this module.  The product package ``voxelizer_b200`` never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
PORT_LIB = os.path.join(HERE, "liboracle_port.so")
REF_LIB = os.path.join(HERE, "_ref", "libvx_ref.so")
REF_GEN_DATA = os.path.join(HERE, "_ref", "ref_gen_data")
REFERENCE_ROOT = "/root/reference"


def build(ref: bool | None = None) -> None:
    """Compile the port; compile oracle/_ref too when the reference tree is present."""
    subprocess.check_call(["make", "-s", "-C", HERE, "port"])
    if ref is None:
        ref = os.path.isdir(os.path.join(REFERENCE_ROOT, "src"))
    if ref:
        subprocess.check_call(["make", "-s", "-C", HERE, "ref", f"REF={REFERENCE_ROOT}"])


def have_ref() -> bool:
    return os.path.exists(REF_LIB) and os.path.exists(REF_GEN_DATA)


class ConfigPod(C.Structure):
    _fields_ = [
        ("voxelSize", C.c_float),
        ("minExtent", C.c_float * 4),
        ("maxExtent", C.c_float * 4),
        ("maxNumScans", C.c_uint32),
        ("priorScans", C.c_uint32),
        ("pastScans", C.c_uint32),
        ("pastDistance", C.c_float),
        ("minRange", C.c_float),
        ("maxRange", C.c_float),
        ("stride_num", C.c_uint32),
        ("stride_distance", C.c_float),
        ("hidecar", C.c_int32),
        ("n_filtered", C.c_uint32),
        ("n_join_pairs", C.c_uint32),
    ]


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _u32(a):
    return np.ascontiguousarray(a, dtype=np.uint32)


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


class Oracle:
    """Uniform view of either checker library (same entry points, different prefix)."""

    def __init__(self, kind: str = "port"):
        self.kind = kind
        if kind == "port":
            if not os.path.exists(PORT_LIB):
                build(ref=False)
            self.lib = C.CDLL(PORT_LIB)
            self.pfx = "orc_"
        elif kind == "ref":
            if not have_ref():
                raise FileNotFoundError("oracle/_ref is not built (needs /root/reference: make -C oracle ref)")
            self.lib = C.CDLL(REF_LIB)
            self.pfx = "vxref_"
        else:
            raise ValueError(kind)
        f = self._fn
        f("last_error").restype = C.c_char_p
        f("config_load").restype = C.c_void_p
        f("config_load").argtypes = [C.c_char_p]
        f("config_free").argtypes = [C.c_void_p]
        f("config_get").argtypes = [C.c_void_p, C.POINTER(ConfigPod)]
        f("config_filtered").argtypes = [C.c_void_p, C.c_void_p]
        f("config_joins").argtypes = [C.c_void_p, C.c_void_p]
        f("grid_new").restype = C.c_void_p
        f("grid_new").argtypes = [C.c_float, C.c_void_p, C.c_void_p]
        f("grid_free").argtypes = [C.c_void_p]
        f("grid_dims").argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        f("grid_clear").argtypes = [C.c_void_p]
        f("grid_insert").argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64]
        f("fill").argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        f("update_occlusions").argtypes = [C.c_void_p]
        f("update_invalid").argtypes = [C.c_void_p, C.c_void_p]
        f("insert_occlusion_labels").argtypes = [C.c_void_p]
        f("occluded_by").restype = C.c_int32
        f("occluded_by").argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p]
        f("grid_dump").restype = C.c_uint64
        f("grid_dump").argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        f("grid_counts").argtypes = [C.c_void_p, C.c_void_p]
        f("grid_hist").restype = C.c_uint64
        f("grid_hist").argtypes = [C.c_void_p, C.c_void_p]
        f("grid_flags").argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        f("grid_save").argtypes = [C.c_void_p, C.c_char_p, C.c_char_p, C.c_char_p]
        f("mat_inverse").argtypes = [C.c_void_p, C.c_void_p]
        f("mat_mul").argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        f("endpoint").argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        f("reader_open").restype = C.c_void_p
        f("reader_open").argtypes = [C.c_char_p]
        f("reader_free").argtypes = [C.c_void_p]
        f("reader_count").restype = C.c_uint32
        f("reader_count").argtypes = [C.c_void_p]
        f("reader_num_poses").restype = C.c_uint32
        f("reader_num_poses").argtypes = [C.c_void_p]
        f("reader_poses").argtypes = [C.c_void_p, C.c_void_p]
        if kind == "port":
            self.lib.orc_gen_data.restype = C.c_int
            self.lib.orc_gen_data.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, C.c_int64, C.c_int64, C.c_void_p, C.c_int]

    def _fn(self, name):
        return getattr(self.lib, self.pfx + name)

    def last_error(self) -> str:
        return (self._fn("last_error")() or b"").decode()

    # ---- config ------------------------------------------------------------------------------
    def load_config(self, path: str) -> "OracleConfig":
        h = self._fn("config_load")(path.encode())
        if not h:
            raise RuntimeError(self.last_error())
        return OracleConfig(self, h)

    # ---- pose algebra --------------------------------------------------------------------------
    def mat_inverse(self, a):
        a = _f32(a).reshape(16)
        out = np.empty(16, np.float32)
        self._fn("mat_inverse")(_ptr(a), _ptr(out))
        return out

    def mat_mul(self, a, b):
        a, b = _f32(a).reshape(16), _f32(b).reshape(16)
        out = np.empty(16, np.float32)
        self._fn("mat_mul")(_ptr(a), _ptr(b), _ptr(out))
        return out

    def endpoint(self, anchor, pose):
        a, p = _f32(anchor).reshape(16), _f32(pose).reshape(16)
        out = np.empty(3, np.float32)
        self._fn("endpoint")(_ptr(a), _ptr(p), _ptr(out))
        return out

    # ---- reader --------------------------------------------------------------------------------
    def read_poses(self, seq_dir: str):
        """(number of .bin files, poses [n,16] column-major) as KittiReader::initialize computes them."""
        h = self._fn("reader_open")(seq_dir.encode())
        if not h:
            raise RuntimeError(self.last_error())
        try:
            n = self._fn("reader_num_poses")(h)
            cnt = self._fn("reader_count")(h)
            out = np.empty((n, 16), np.float32)
            self._fn("reader_poses")(h, _ptr(out))
            return cnt, out
        finally:
            self._fn("reader_free")(h)

    def grid(self, resolution, min_extent, max_extent) -> "OracleGrid":
        return OracleGrid(self, resolution, min_extent, max_extent)

    # ---- whole program -------------------------------------------------------------------------
    def gen_data(self, cfg: str, seq_dir: str, out_dir: str, first_frame: int = 0, max_frames: int = -1):
        """Runs the gen_data program.  port: in-process (returns stage stats); ref: the reference binary."""
        if self.kind == "port":
            stats = np.zeros(6, np.float64)
            rc = self.lib.orc_gen_data(cfg.encode(), seq_dir.encode(), out_dir.encode(), first_frame, max_frames, _ptr(stats), 1)
            if rc != 0:
                raise RuntimeError(self.last_error())
            return dict(fill_s=stats[0], occl_s=stats[1], invalid_s=stats[2], save_s=stats[3], frames=int(stats[4]), skipped=int(stats[5]))
        if first_frame != 0 or max_frames != -1:
            raise ValueError("the reference binary always processes the whole sequence")
        subprocess.check_call([REF_GEN_DATA, cfg, seq_dir, out_dir], stdout=subprocess.DEVNULL)
        return None


class OracleConfig:
    def __init__(self, o: Oracle, handle):
        self.o, self.h = o, handle
        pod = ConfigPod()
        o._fn("config_get")(handle, C.byref(pod))
        self.pod = pod
        self.filtered = np.zeros(pod.n_filtered, np.uint32)
        if pod.n_filtered:
            o._fn("config_filtered")(handle, _ptr(self.filtered))
        self.join_pairs = np.zeros((pod.n_join_pairs, 2), np.uint32)  # (key, member)
        if pod.n_join_pairs:
            o._fn("config_joins")(handle, _ptr(self.join_pairs))

    def as_dict(self):
        p = self.pod
        return dict(
            voxelSize=p.voxelSize, minExtent=list(p.minExtent), maxExtent=list(p.maxExtent), maxNumScans=p.maxNumScans,
            priorScans=p.priorScans, pastScans=p.pastScans, pastDistance=p.pastDistance, minRange=p.minRange,
            maxRange=p.maxRange, stride_num=p.stride_num, stride_distance=p.stride_distance, hidecar=p.hidecar,
            filtered=self.filtered.tolist(), join_pairs=self.join_pairs.tolist(),
        )

    def __del__(self):
        try:
            self.o._fn("config_free")(self.h)
        except Exception:
            pass


class OracleGrid:
    """One VoxelGrid.  Arrays come back in the reference-internal index order i + j*Sx + k*Sx*Sy unless
    stated otherwise; ``to_output_order`` converts to the file order x*Ny*Nz + y*Nz + z."""

    def __init__(self, o: Oracle, resolution, min_extent, max_extent):
        self.o = o
        mn = _f32(list(min_extent)[:3] + [1.0])
        mx = _f32(list(max_extent)[:3] + [1.0])
        self.h = o._fn("grid_new")(C.c_float(resolution), _ptr(mn), _ptr(mx))
        size = np.zeros(3, np.uint32)
        off = np.zeros(4, np.float32)
        res = C.c_float()
        o._fn("grid_dims")(self.h, _ptr(size), _ptr(off), C.byref(res))
        self.size = tuple(int(x) for x in size)
        self.offset = off[:3].copy()
        self.resolution = np.float32(res.value)
        self.nvox = self.size[0] * self.size[1] * self.size[2]

    def __del__(self):
        try:
            self.o._fn("grid_free")(self.h)
        except Exception:
            pass

    def clear(self):
        self.o._fn("grid_clear")(self.h)

    def insert(self, p4, labels):
        p4, labels = _f32(p4), _u32(labels)
        self.o._fn("grid_insert")(self.h, _ptr(p4), _ptr(labels), len(labels))

    def fill(self, cfg: OracleConfig, anchor, scans, labels, poses):
        """fillVoxelGrid: scans = list of [n,4] float32, labels = list of [n] uint32, poses = [k,16]."""
        k = len(scans)
        scans = [_f32(s).reshape(-1, 4) for s in scans]
        labels = [_u32(l) for l in labels]
        pa = (C.c_void_p * k)(*[s.ctypes.data for s in scans])
        la = (C.c_void_p * k)(*[l.ctypes.data for l in labels])
        counts = _u32([len(l) for l in labels])
        poses = _f32(poses).reshape(k, 16)
        anchor = _f32(anchor).reshape(16)
        self.o._fn("fill")(self.h, cfg.h, _ptr(anchor), k, pa, la, _ptr(counts), _ptr(poses))

    def update_occlusions(self):
        self.o._fn("update_occlusions")(self.h)

    def update_invalid(self, endpoint):
        e = _f32(endpoint).reshape(3)
        self.o._fn("update_invalid")(self.h, _ptr(e))

    def insert_occlusion_labels(self):
        """VoxelGrid::insertOcclusionLabels (VoxelGrid.cpp:75-96)."""
        self.o._fn("insert_occlusion_labels")(self.h)

    def occluded_by(self, i: int, j: int, k: int, endpoint, capacity: int = 4096):
        """VoxelGrid::occludedBy(i, j, k, endpoint, &visited): (occluder internal index or -1, visited [n,3])."""
        e = _f32(endpoint).reshape(3)
        vis = np.zeros((capacity, 3), np.int32)
        n = C.c_uint32()
        r = self.o._fn("occluded_by")(self.h, i, j, k, _ptr(e), _ptr(vis), capacity, C.byref(n))
        return int(r), vis[: min(int(n.value), capacity)].copy()

    def dump(self, what: str):
        code = {"occlusions": 0, "invalid": 1, "memo": 2}[what]
        n = self.o._fn("grid_dump")(self.h, code, None)
        out = np.empty(n, np.int32)
        if n:
            self.o._fn("grid_dump")(self.h, code, _ptr(out))
        return out

    def counts(self):
        out = np.empty(self.nvox, np.uint32)
        self.o._fn("grid_counts")(self.h, _ptr(out))
        return out

    def hist(self):
        """[m,3] (internal voxel index, label, count) sorted by (voxel, label)."""
        n = self.o._fn("grid_hist")(self.h, None)
        out = np.empty((n, 3), np.uint32)
        if n:
            self.o._fn("grid_hist")(self.h, _ptr(out))
        return out

    def flags(self):
        """(occluded, invalid) as uint8 0/1 arrays in OUTPUT order."""
        occ = np.empty(self.nvox, np.uint8)
        inv = np.empty(self.nvox, np.uint8)
        self.o._fn("grid_flags")(self.h, _ptr(occ), _ptr(inv))
        return occ, inv

    def save(self, directory: str, basename: str, mode: str):
        self.o._fn("grid_save")(self.h, directory.encode(), basename.encode(), mode.encode())

    def to_output_order(self, internal):
        sx, sy, sz = self.size
        return np.ascontiguousarray(np.asarray(internal).reshape(sz, sy, sx).transpose(2, 1, 0)).reshape(-1)


def pack_msb(bits) -> np.ndarray:
    """voxelize_utils.cpp:205-216 -- element 8n in bit 7 of byte n."""
    b = (np.asarray(bits) > 0).astype(np.uint8)
    return np.packbits(b[: (len(b) // 8) * 8], bitorder="big")
