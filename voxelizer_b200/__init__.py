"""voxelizer_b200 -- B200-native backend for jbehley/voxelizer's gen_data hot path.

Python is only a binding here: the product is the C ABI in ``include/voxelizer_b200.h`` (CUDA kernels
for sm_100a in ``csrc/``) plus the C++ host layer in ``host/`` that mirrors the reference's own API.
This module loads the in-tree shared libraries with ctypes.  If the CUDA library is missing it raises --
there is no CPU fallback, and the CPU checker used by the tests is never imported from here.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass

import numpy as np

PKG = os.path.dirname(os.path.abspath(__file__))
LIB_DIR = os.path.join(PKG, "lib")
# developer switch for same-box A/B timing of two builds of the kernels (tools/dev_gpu_perf.py); never set in tests / bench
CUDA_LIB = os.environ.get("VX_B200_CUDA_LIB") or os.path.join(LIB_DIR, "libvoxelizer_b200.so")
HOST_LIB = os.path.join(LIB_DIR, "libvxhost.so")
SYNTH_LIB = os.path.join(LIB_DIR, "libvxsynth.so")
GEN_DATA_EXE = os.path.join(PKG, "bin", "vx_gen_data")

# every symbol include/voxelizer_b200.h declares
C_ABI_SYMBOLS = (
    "vx_last_global_error", "vx_last_error", "vx_device_count", "vx_create", "vx_destroy", "vx_grid_info_get",
    "vx_frames_in_flight", "vx_batches_in_flight", "vx_host_alloc", "vx_host_free", "vx_upload_scan", "vx_upload_ticket", "vx_upload_wait",
    "vx_evict_scan", "vx_evict_all",
    "vx_process_frames", "vx_submit_frames", "vx_wait", "vx_sync", "vx_stage_times_get", "vx_stage_times_reset",
    "vx_debug_dump", "vx_trace_ray",
)

VX_OK = 0
DUMP_HIST_TARGET, DUMP_HIST_INPUT, DUMP_OCCUPANCY, DUMP_OCCLUDED, DUMP_INVALID, DUMP_FILL_SOURCE = range(6)
FRAME_INSERT_OCCLUSION_LABELS, FRAME_INPUT_IS_TARGET = 1, 2


class VxError(RuntimeError):
    pass


class GridDesc(C.Structure):
    _fields_ = [("resolution", C.c_float), ("min_extent", C.c_float * 3), ("max_extent", C.c_float * 3)]


class FilterDesc(C.Structure):
    _fields_ = [
        ("min_range", C.c_float), ("max_range", C.c_float), ("hidecar", C.c_int32),
        ("filtered_labels", C.c_void_p), ("n_filtered", C.c_uint32),
        ("join_pairs", C.c_void_p), ("n_join_pairs", C.c_uint32),
    ]


class Options(C.Structure):
    _fields_ = [("frames_in_flight", C.c_uint32), ("table_log2_capacity", C.c_uint32), ("stream", C.c_void_p),
                ("debug", C.c_uint32), ("table_sets", C.c_uint32), ("vis_mode", C.c_uint32), ("occlusion_labels", C.c_uint32)]


class FrameDesc(C.Structure):
    _fields_ = [
        ("n_prior", C.c_uint32), ("prior_ids", C.c_void_p), ("ap_prior", C.c_void_p),
        ("n_past", C.c_uint32), ("past_ids", C.c_void_p), ("ap_past", C.c_void_p),
        ("n_endpoints", C.c_uint32), ("endpoints", C.c_void_p),
        ("out_bin", C.c_void_p), ("out_label", C.c_void_p), ("out_occluded", C.c_void_p), ("out_invalid", C.c_void_p),
        ("flags", C.c_uint32), ("reserved", C.c_uint32),
    ]


class GridInfo(C.Structure):
    _fields_ = [
        ("size", C.c_uint32 * 3), ("offset", C.c_float * 3), ("resolution", C.c_float),
        ("num_voxels", C.c_uint64), ("packed_bytes", C.c_uint64), ("label_bytes", C.c_uint64),
    ]


class StageTimes(C.Structure):
    _fields_ = [
        ("fill_ms", C.c_double), ("vote_ms", C.c_double), ("emit_ms", C.c_double), ("visibility_ms", C.c_double),
        ("pack_ms", C.c_double), ("total_ms", C.c_double), ("frames", C.c_uint64), ("points_read", C.c_uint64),
        ("fill_launches", C.c_uint64), ("vote_launches", C.c_uint64), ("emit_launches", C.c_uint64),
        ("visibility_launches", C.c_uint64), ("pack_launches", C.c_uint64), ("rays_cast", C.c_uint64), ("dda_steps", C.c_uint64),
        ("vis_waves", C.c_uint64), ("vis_resolve_waves", C.c_uint64), ("vis_rounds", C.c_uint64), ("vis_cycles", C.c_uint64 * 8),
    ]

    def as_dict(self):
        return {n: (list(getattr(self, n)) if n == "vis_cycles" else getattr(self, n)) for n, _ in self._fields_}


class ConfigPod(C.Structure):
    _fields_ = [
        ("voxelSize", C.c_float), ("minExtent", C.c_float * 4), ("maxExtent", C.c_float * 4),
        ("maxNumScans", C.c_uint32), ("priorScans", C.c_uint32), ("pastScans", C.c_uint32),
        ("pastDistance", C.c_float), ("minRange", C.c_float), ("maxRange", C.c_float),
        ("stride_num", C.c_uint32), ("stride_distance", C.c_float), ("hidecar", C.c_int32),
        ("n_filtered", C.c_uint32), ("n_join_pairs", C.c_uint32),
    ]


class GenDataStats(C.Structure):
    _fields_ = [
        ("frames", C.c_uint64), ("skipped", C.c_uint64), ("wall_s", C.c_double), ("read_s", C.c_double),
        ("gpu_s", C.c_double), ("write_s", C.c_double), ("stages", StageTimes),
    ]


class SynthParams(C.Structure):
    _fields_ = [("seed", C.c_uint64), ("n_beams", C.c_uint32), ("n_azimuth", C.c_uint32)] + [
        (n, C.c_float)
        for n in ("elev_top_deg", "elev_bottom_deg", "max_return_m", "sensor_height_m", "speed_m", "noise_m",
                  "unlabeled_frac", "outlier_frac", "sway_amp_m", "yaw_amp_rad")
    ]


_cuda = None
_host = None
_synth = None


def _need(path: str, what: str):
    if not os.path.exists(path):
        raise VxError(f"{what} is not built: {path} missing. Run `python -m voxelizer_b200.build` "
                      "(or __graft_entry__.build()). There is no CPU fallback.")
    return C.CDLL(path, mode=C.RTLD_GLOBAL)


def cuda_lib():
    """The CUDA library (C ABI).  Loading it does not need a GPU; creating a context does."""
    global _cuda
    if _cuda is None:
        L = _need(CUDA_LIB, "the CUDA extension")
        L.vx_last_global_error.restype = C.c_char_p
        L.vx_last_error.restype = C.c_char_p
        L.vx_last_error.argtypes = [C.c_void_p]
        L.vx_create.argtypes = [C.c_int, C.POINTER(GridDesc), C.POINTER(FilterDesc), C.POINTER(Options), C.POINTER(C.c_void_p)]
        L.vx_destroy.argtypes = [C.c_void_p]
        L.vx_grid_info_get.argtypes = [C.c_void_p, C.POINTER(GridInfo)]
        L.vx_frames_in_flight.restype = C.c_uint32
        L.vx_frames_in_flight.argtypes = [C.c_void_p]
        L.vx_batches_in_flight.restype = C.c_uint32
        L.vx_batches_in_flight.argtypes = [C.c_void_p]
        L.vx_host_alloc.restype = C.c_void_p
        L.vx_host_alloc.argtypes = [C.c_size_t]
        L.vx_host_free.argtypes = [C.c_void_p]
        L.vx_upload_scan.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_uint32]
        L.vx_upload_ticket.restype = C.c_uint64
        L.vx_upload_ticket.argtypes = [C.c_void_p]
        L.vx_upload_wait.argtypes = [C.c_void_p, C.c_uint64]
        L.vx_evict_scan.argtypes = [C.c_void_p, C.c_uint32]
        L.vx_evict_all.argtypes = [C.c_void_p]
        L.vx_process_frames.argtypes = [C.c_void_p, C.POINTER(FrameDesc), C.c_uint32]
        L.vx_submit_frames.argtypes = [C.c_void_p, C.POINTER(FrameDesc), C.c_uint32, C.POINTER(C.c_uint64)]
        L.vx_wait.argtypes = [C.c_void_p, C.c_uint64]
        L.vx_trace_ray.argtypes = [C.c_void_p, C.c_uint32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_uint32,
                                   C.POINTER(C.c_uint32), C.POINTER(C.c_int32)]
        L.vx_sync.argtypes = [C.c_void_p]
        L.vx_stage_times_get.argtypes = [C.c_void_p, C.POINTER(StageTimes)]
        L.vx_stage_times_reset.argtypes = [C.c_void_p]
        L.vx_debug_dump.argtypes = [C.c_void_p, C.c_uint32, C.c_int, C.c_void_p, C.POINTER(C.c_uint64)]
        _cuda = L
    return _cuda


def host_lib():
    global _host
    if _host is None:
        cuda_lib()
        L = _need(HOST_LIB, "the host library")
        L.vxh_last_error.restype = C.c_char_p
        L.vxh_config_load.restype = C.c_void_p
        L.vxh_config_load.argtypes = [C.c_char_p]
        L.vxh_config_free.argtypes = [C.c_void_p]
        L.vxh_config_get.argtypes = [C.c_void_p, C.POINTER(ConfigPod)]
        L.vxh_config_filtered.argtypes = [C.c_void_p, C.c_void_p]
        L.vxh_config_joins.argtypes = [C.c_void_p, C.c_void_p]
        L.vxh_mat_inverse.argtypes = [C.c_void_p, C.c_void_p]
        L.vxh_mat_mul.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.vxh_frame_transforms.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p]
        L.vxh_reader_open.restype = C.c_void_p
        L.vxh_reader_open.argtypes = [C.c_char_p]
        L.vxh_reader_free.argtypes = [C.c_void_p]
        L.vxh_reader_count.restype = C.c_uint32
        L.vxh_reader_count.argtypes = [C.c_void_p]
        L.vxh_reader_num_poses.restype = C.c_uint32
        L.vxh_reader_num_poses.argtypes = [C.c_void_p]
        L.vxh_reader_poses.argtypes = [C.c_void_p, C.c_void_p]
        L.vxh_reader_window.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p]
        L.vxh_reader_load.argtypes = [C.c_void_p, C.c_uint32, C.c_int, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p]
        L.vxh_plan_targets.restype = C.c_uint32
        L.vxh_plan_targets.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32]
        L.vxh_gen_data.restype = C.c_int
        L.vxh_gen_data.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, C.c_int, C.c_uint32, C.c_uint32, C.c_uint32,
                                   C.c_int64, C.c_int64, C.c_int, C.c_int, C.POINTER(GenDataStats)]
        _host = L
    return _host


def synth_lib():
    global _synth
    if _synth is None:
        L = _need(SYNTH_LIB, "the synthetic sequence generator")
        L.vxs_default_params.argtypes = [C.POINTER(SynthParams)]
        L.vxs_scan.restype = C.c_uint32
        L.vxs_scan.argtypes = [C.POINTER(SynthParams), C.c_uint32, C.c_void_p, C.c_void_p, C.c_uint32]
        L.vxs_velo_pose.argtypes = [C.POINTER(SynthParams), C.c_uint32, C.c_void_p]
        L.vxs_calib_tr.argtypes = [C.c_void_p]
        L.vxs_generate_many.restype = C.c_int
        L.vxs_generate_many.argtypes = [C.POINTER(SynthParams), C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_int]
        L.vxs_write_sequence.restype = C.c_int
        L.vxs_write_sequence.argtypes = [C.POINTER(SynthParams), C.c_char_p, C.c_uint32, C.c_int, C.c_int]
        _synth = L
    return _synth


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _u32(a):
    return np.ascontiguousarray(a, dtype=np.uint32)


# ------------------------------------------------------------------------------------------------------
# Host layer: Config / KittiReader / planning (mirrors the reference's host API)
# ------------------------------------------------------------------------------------------------------
class Config:
    """parseConfiguration(filename) of the host layer (voxelizer_b200/host/config.cpp)."""

    def __init__(self, path: str):
        H = host_lib()
        self._h = H.vxh_config_load(path.encode())
        if not self._h:
            raise VxError(H.vxh_last_error().decode())
        self.path = path
        self.pod = ConfigPod()
        H.vxh_config_get(self._h, C.byref(self.pod))
        self.filtered = np.zeros(self.pod.n_filtered, np.uint32)
        if self.pod.n_filtered:
            H.vxh_config_filtered(self._h, _p(self.filtered))
        self.join_pairs = np.zeros((self.pod.n_join_pairs, 2), np.uint32)
        if self.pod.n_join_pairs:
            H.vxh_config_joins(self._h, _p(self.join_pairs))

    def __del__(self):
        try:
            host_lib().vxh_config_free(self._h)
        except Exception:
            pass

    def as_dict(self):
        p = self.pod
        return dict(
            voxelSize=p.voxelSize, minExtent=list(p.minExtent), maxExtent=list(p.maxExtent), maxNumScans=p.maxNumScans,
            priorScans=p.priorScans, pastScans=p.pastScans, pastDistance=p.pastDistance, minRange=p.minRange,
            maxRange=p.maxRange, stride_num=p.stride_num, stride_distance=p.stride_distance, hidecar=p.hidecar,
            filtered=self.filtered.tolist(), join_pairs=self.join_pairs.tolist(),
        )


def mat_inverse(a):
    a = _f32(a).reshape(16)
    out = np.empty(16, np.float32)
    host_lib().vxh_mat_inverse(_p(a), _p(out))
    return out


def mat_mul(a, b):
    a, b = _f32(a).reshape(16), _f32(b).reshape(16)
    out = np.empty(16, np.float32)
    host_lib().vxh_mat_mul(_p(a), _p(b), _p(out))
    return out


def frame_transforms(anchor, poses):
    """(ap [n,16], endpoints [n,3]) for poses [n,16] relative to the anchor pose (column-major)."""
    anchor = _f32(anchor).reshape(16)
    poses = _f32(poses).reshape(-1, 16)
    n = len(poses)
    ap = np.empty((n, 16), np.float32)
    ep = np.empty((n, 3), np.float32)
    if n:
        host_lib().vxh_frame_transforms(_p(anchor), _p(poses), n, _p(ap), _p(ep))
    return ap, ep


def read_poses(seq_dir: str):
    """(number of velodyne files, poses [n,16]) exactly as the host KittiReader computes them."""
    H = host_lib()
    h = H.vxh_reader_open(seq_dir.encode())
    if not h:
        raise VxError(H.vxh_last_error().decode())
    try:
        n = H.vxh_reader_num_poses(h)
        out = np.empty((n, 16), np.float32)
        if n:
            H.vxh_reader_poses(h, _p(out))
        return H.vxh_reader_count(h), out
    finally:
        H.vxh_reader_free(h)


def read_scan(seq_dir: str, t: int, direct: bool, capacity: int = 1 << 20):
    """Scan t through the host KittiReader: direct=True is the pipeline's loader (raw labels into caller memory),
    False the reference-shaped one (labels & 0xFFFF).  Returns (xyzr [n,4], labels [n], labelled) or None if the scan
    does not fit `capacity` points."""
    H = host_lib()
    h = H.vxh_reader_open(seq_dir.encode())
    if not h:
        raise VxError(H.vxh_last_error().decode())
    try:
        xyzr = np.zeros((capacity, 4), np.float32)
        labels = np.zeros(capacity, np.uint32)
        n, labelled = C.c_uint32(0), C.c_uint32(0)
        rc = H.vxh_reader_load(h, t, int(direct), _p(xyzr), _p(labels), capacity, C.byref(n), C.byref(labelled))
        if rc < 0:
            raise VxError(H.vxh_last_error().decode())
        if rc == 1:
            return None
        return xyzr[: n.value].copy(), labels[: n.value].copy(), int(labelled.value)
    finally:
        H.vxh_reader_free(h)


def window(scan_count: int, prior: int, past: int, idx: int):
    """Scan index ranges (prior_begin, prior_end, past_begin, past_end) of KittiReader::retrieve(idx)."""
    pb = max(0, idx - prior)
    pe = idx + 1
    qb = min(scan_count - 1, idx + 1)
    qe = max(qb, min(scan_count, idx + past + 1))
    return pb, pe, qb, qe


def plan_targets(cfg: Config, scan_count: int, poses):
    poses = _f32(poses).reshape(-1, 16)
    out = np.empty(max(1, scan_count), np.uint32)
    n = host_lib().vxh_plan_targets(cfg._h, scan_count, _p(poses), len(poses), _p(out), len(out))
    return out[:n].copy()


def gen_data(cfg_path: str, seq_dir: str, out_dir: str, device: int = 0, shard=(0, 1), frames_in_flight: int = 0,
             first_frame: int = 0, max_frames: int = -1, write_files: bool = True, verbose: bool = False) -> dict:
    """The whole gen_data program on one GPU (host/gen_data_pipeline.cpp)."""
    H = host_lib()
    st = GenDataStats()
    rc = H.vxh_gen_data(cfg_path.encode(), seq_dir.encode(), out_dir.encode(), device, shard[0], shard[1], frames_in_flight,
                        first_frame, max_frames, int(write_files), int(verbose), C.byref(st))
    if rc != 0:
        raise VxError(H.vxh_last_error().decode())
    d = dict(frames=st.frames, skipped=st.skipped, wall_s=st.wall_s, read_s=st.read_s, gpu_s=st.gpu_s, write_s=st.write_s)
    d["stages"] = st.stages.as_dict()
    return d


# ------------------------------------------------------------------------------------------------------
# The CUDA backend through the C ABI
# ------------------------------------------------------------------------------------------------------
@dataclass
class FrameSpec:
    prior_ids: np.ndarray
    ap_prior: np.ndarray
    past_ids: np.ndarray
    ap_past: np.ndarray
    endpoints: np.ndarray
    flags: int = 0


class PinnedBuffer:
    """Page-locked host memory viewed as a numpy array."""

    def __init__(self, nbytes: int):
        L = cuda_lib()
        self.nbytes = int(nbytes)
        self.ptr = L.vx_host_alloc(self.nbytes)
        if not self.ptr:
            raise VxError("vx_host_alloc failed")
        self.array = np.ctypeslib.as_array((C.c_uint8 * self.nbytes).from_address(self.ptr))

    def free(self):
        if self.ptr:
            self.array = None
            cuda_lib().vx_host_free(self.ptr)
            self.ptr = None


class Backend:
    """One vx_ctx (one GPU)."""

    def __init__(self, resolution, min_extent, max_extent, min_range, max_range, hidecar=True, filtered=(), join_pairs=(),
                 device: int = 0, frames_in_flight: int = 0, table_log2_capacity: int = 0, debug: bool = False, stream: int = 0,
                 table_sets: int = 0, vis_mode: int = 0, occlusion_labels: bool = False):
        L = cuda_lib()
        gd = GridDesc(resolution=float(resolution))
        for a in range(3):
            gd.min_extent[a] = float(min_extent[a])
            gd.max_extent[a] = float(max_extent[a])
        self._filtered = _u32(np.asarray(filtered, dtype=np.uint32).reshape(-1))
        self._joins = _u32(np.asarray(join_pairs, dtype=np.uint32).reshape(-1))
        fd = FilterDesc(min_range=float(min_range), max_range=float(max_range), hidecar=int(bool(hidecar)),
                        filtered_labels=self._filtered.ctypes.data if len(self._filtered) else None, n_filtered=len(self._filtered),
                        join_pairs=self._joins.ctypes.data if len(self._joins) else None, n_join_pairs=len(self._joins) // 2)
        op = Options(frames_in_flight=frames_in_flight, table_log2_capacity=table_log2_capacity, stream=stream or None,
                     debug=1 if debug else 0, table_sets=table_sets, vis_mode=vis_mode,
                     occlusion_labels=1 if occlusion_labels else 0)
        ctx = C.c_void_p()
        rc = L.vx_create(device, C.byref(gd), C.byref(fd), C.byref(op), C.byref(ctx))
        if rc != VX_OK:
            raise VxError(f"vx_create failed ({rc}): {L.vx_last_global_error().decode()}")
        self._ctx = ctx
        gi = GridInfo()
        L.vx_grid_info_get(self._ctx, C.byref(gi))
        self.size = tuple(gi.size)
        self.offset = np.array(list(gi.offset), np.float32)
        self.resolution = np.float32(gi.resolution)
        self.nvox = int(gi.num_voxels)
        self.packed_bytes = int(gi.packed_bytes)
        self.label_bytes = int(gi.label_bytes)
        self.frames_in_flight = int(L.vx_frames_in_flight(self._ctx))
        self.batches_in_flight = int(L.vx_batches_in_flight(self._ctx))

    @classmethod
    def from_config(cls, cfg: Config, **kw):
        p = cfg.pod
        return cls(p.voxelSize, list(p.minExtent)[:3], list(p.maxExtent)[:3], p.minRange, p.maxRange, bool(p.hidecar),
                   cfg.filtered, cfg.join_pairs, **kw)

    def close(self):
        if self._ctx:
            cuda_lib().vx_destroy(self._ctx)
            self._ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc, what):
        if rc != VX_OK:
            raise VxError(f"{what} failed ({rc}): {cuda_lib().vx_last_error(self._ctx).decode()}")

    def upload_scan(self, scan_id: int, xyzr, labels):
        xyzr = _f32(xyzr).reshape(-1, 4)
        labels = _u32(labels).reshape(-1)
        assert len(xyzr) == len(labels)
        self._check(cuda_lib().vx_upload_scan(self._ctx, scan_id, _p(xyzr), _p(labels), len(labels)), "vx_upload_scan")

    def upload_scan_ptr(self, scan_id: int, xyzr_ptr: int, labels_ptr: int, n: int):
        self._check(cuda_lib().vx_upload_scan(self._ctx, scan_id, xyzr_ptr, labels_ptr, n), "vx_upload_scan")

    def evict_all(self):
        cuda_lib().vx_evict_all(self._ctx)

    def build_descs(self, frames, outputs=None):
        """ctypes array of vx_frame_desc for a list of FrameSpec.  outputs: optional list of dicts with numpy
        arrays under 'bin', 'label', 'occluded', 'invalid'.  Returns (array, keepalive)."""
        n = len(frames)
        arr = (FrameDesc * n)()
        keep = []
        for i, f in enumerate(frames):
            pid, pap = _u32(f.prior_ids), _f32(f.ap_prior).reshape(-1, 16)
            qid, qap = _u32(f.past_ids), _f32(f.ap_past).reshape(-1, 16)
            ep = _f32(f.endpoints).reshape(-1, 3)
            keep += [pid, pap, qid, qap, ep]
            d = arr[i]
            d.n_prior, d.prior_ids, d.ap_prior = len(pid), pid.ctypes.data, pap.ctypes.data
            d.n_past, d.past_ids, d.ap_past = len(qid), qid.ctypes.data, qap.ctypes.data
            d.n_endpoints, d.endpoints = len(ep), ep.ctypes.data
            d.flags = int(getattr(f, "flags", 0))
            if outputs is not None:
                o = outputs[i]
                d.out_bin, d.out_label = o["bin"].ctypes.data, o["label"].ctypes.data
                d.out_occluded, d.out_invalid = o["occluded"].ctypes.data, o["invalid"].ctypes.data
        return arr, keep

    def new_outputs(self, n: int):
        return [dict(bin=np.zeros(self.packed_bytes, np.uint8), label=np.zeros(self.nvox, np.uint16),
                     occluded=np.zeros(self.packed_bytes, np.uint8), invalid=np.zeros(self.packed_bytes, np.uint8)) for _ in range(n)]

    def process(self, frames, outputs=None):
        """Runs frames; returns the outputs (allocated if not given)."""
        if outputs is None:
            outputs = self.new_outputs(len(frames))
        arr, keep = self.build_descs(frames, outputs)
        self._check(cuda_lib().vx_process_frames(self._ctx, arr, len(frames)), "vx_process_frames")
        del keep
        return outputs

    def process_descs(self, arr, n):
        self._check(cuda_lib().vx_process_frames(self._ctx, arr, n), "vx_process_frames")

    def submit_descs(self, arr, n) -> int:
        """Asynchronous: returns a ticket for wait(); scans with other ids may be uploaded meanwhile."""
        t = C.c_uint64()
        self._check(cuda_lib().vx_submit_frames(self._ctx, arr, n, C.byref(t)), "vx_submit_frames")
        return int(t.value)

    def wait(self, ticket: int = 0):
        self._check(cuda_lib().vx_wait(self._ctx, ticket), "vx_wait")

    def sync(self):
        self._check(cuda_lib().vx_sync(self._ctx), "vx_sync")

    def evict_scan(self, scan_id: int):
        cuda_lib().vx_evict_scan(self._ctx, scan_id)

    def trace_ray(self, frame_index: int, i: int, j: int, k: int, endpoint, capacity: int = 4096):
        """VoxelGrid::occludedBy(i, j, k, endpoint, &visited) on frame `frame_index` of the last job:
        (occluder as reference-internal index or -1, visited [n,3] int32)."""
        e = _f32(endpoint).reshape(3)
        vis = np.zeros((capacity, 3), np.int32)
        n, occ = C.c_uint32(), C.c_int32()
        self._check(cuda_lib().vx_trace_ray(self._ctx, frame_index, i, j, k, _p(e), _p(vis), capacity, C.byref(n), C.byref(occ)), "vx_trace_ray")
        return int(occ.value), vis[: min(int(n.value), capacity)].copy()

    def stage_times(self) -> dict:
        st = StageTimes()
        cuda_lib().vx_stage_times_get(self._ctx, C.byref(st))
        return st.as_dict()

    def reset_stage_times(self):
        cuda_lib().vx_stage_times_reset(self._ctx)

    def dump(self, frame_index: int, kind: int):
        L = cuda_lib()
        cnt = C.c_uint64()
        self._check(L.vx_debug_dump(self._ctx, frame_index, kind, None, C.byref(cnt)), "vx_debug_dump")
        if kind in (DUMP_HIST_TARGET, DUMP_HIST_INPUT):
            out = np.zeros((cnt.value, 3), np.uint32)
        elif kind == DUMP_FILL_SOURCE:
            out = np.zeros(cnt.value, np.uint32)
        else:
            out = np.zeros(cnt.value, np.uint8)
        if cnt.value:
            self._check(L.vx_debug_dump(self._ctx, frame_index, kind, _p(out), C.byref(cnt)), "vx_debug_dump")
        return out


# ------------------------------------------------------------------------------------------------------
# Synthetic data (a data source for product and oracle alike)
# ------------------------------------------------------------------------------------------------------
def synth_params(**kw) -> SynthParams:
    p = SynthParams()
    synth_lib().vxs_default_params(C.byref(p))
    for k, v in kw.items():
        setattr(p, k, v)
    return p


def synth_scan(params: SynthParams, t: int, out_xyzr=None, out_labels=None):
    """Scan t as (xyzr [n,4] float32, labels [n] uint32 with instance ids in the upper 16 bits)."""
    cap = params.n_beams * params.n_azimuth
    xyzr = out_xyzr if out_xyzr is not None else np.empty((cap, 4), np.float32)
    labels = out_labels if out_labels is not None else np.empty(cap, np.uint32)
    n = synth_lib().vxs_scan(C.byref(params), t, _p(xyzr), _p(labels), cap)
    return xyzr[:n], labels[:n]


def synth_generate_many(params: SynthParams, t0: int, n: int, xyzr_ptr: int, labels_ptr: int, stride_points: int, threads: int = 0):
    """Scans t0 .. t0+n-1 into caller memory (scan s at point offset s*stride_points); returns counts [n]."""
    if threads <= 0:
        threads = min(64, os.cpu_count() or 1)
    counts = np.zeros(n, np.uint32)
    rc = synth_lib().vxs_generate_many(C.byref(params), t0, n, xyzr_ptr, labels_ptr, stride_points, _p(counts), threads)
    if rc != 0:
        raise VxError("vxs_generate_many failed")
    return counts


def synth_velo_poses(params: SynthParams, n: int):
    """Intended velodyne poses, [n,16] column-major float64."""
    out = np.empty((n, 16), np.float64)
    tmp = np.empty(16, np.float64)
    for t in range(n):
        synth_lib().vxs_velo_pose(C.byref(params), t, _p(tmp))
        out[t] = tmp.reshape(4, 4).T.reshape(16)
    return out


def write_sequence(params: SynthParams, directory: str, n_scans: int, write_labels: bool = True, threads: int = 0):
    if threads <= 0:
        threads = min(32, os.cpu_count() or 1)
    rc = synth_lib().vxs_write_sequence(C.byref(params), directory.encode(), n_scans, int(write_labels), threads)
    if rc != 0:
        raise VxError(f"vxs_write_sequence failed ({rc})")
