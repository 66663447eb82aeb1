"""Developer diagnostic: where does the time of uploading a sequence go?  (1) one big H2D copy from pinned memory, (2) the same bytes in
scan-sized chunks into one preallocated buffer, (3) vx_upload_scan into a fresh context (stream-ordered pool allocations that need new
memory), (4) evict + vx_upload_scan again (the pool reuses its blocks)."""
import ctypes as C
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import voxelizer_b200 as vx  # noqa: E402
from util import cfg_path  # noqa: E402

n, CAP = int(sys.argv[1]) if len(sys.argv) > 1 else 1500, 124000
rt = C.CDLL("libcudart.so.12")
pin = vx.PinnedBuffer(n * CAP * 20)
pin.array[:] = 1
total = n * CAP * 20
dev = C.c_void_p()
assert rt.cudaMalloc(C.byref(dev), C.c_size_t(total)) == 0
st = C.c_void_p()
assert rt.cudaStreamCreate(C.byref(st)) == 0
H2D = 1


def timed(fn, what):
    rt.cudaDeviceSynchronize()
    t0 = time.perf_counter()
    fn()
    rt.cudaDeviceSynchronize()
    dt = time.perf_counter() - t0
    print(f"{what}: {dt * 1e3:7.1f} ms = {total / dt / 1e9:6.1f} GB/s", flush=True)


def big():
    rt.cudaMemcpyAsync(dev, C.c_void_p(pin.ptr), C.c_size_t(total), H2D, st)


def chunks():
    off = 0
    for _ in range(n):
        rt.cudaMemcpyAsync(C.c_void_p(dev.value + off), C.c_void_p(pin.ptr + off), C.c_size_t(CAP * 16), H2D, st)
        off += CAP * 16
        rt.cudaMemcpyAsync(C.c_void_p(dev.value + off), C.c_void_p(pin.ptr + off), C.c_size_t(CAP * 4), H2D, st)
        off += CAP * 4


timed(big, "warm-up, one copy        ")
timed(big, "one copy                 ")
timed(chunks, "scan-sized chunks        ")
rt.cudaFree(dev)
be = vx.Backend.from_config(vx.Config(cfg_path("semantic_kitti")))


def uploads():
    for t in range(n):
        be.upload_scan_ptr(t, pin.ptr + t * CAP * 20, pin.ptr + t * CAP * 20 + CAP * 16, CAP)
    be.sync()


def reuploads():
    for t in range(n):
        be.evict_scan(t)
    uploads()


timed(uploads, "vx_upload_scan, new pool ")
timed(reuploads, "evict + vx_upload_scan   ")
timed(reuploads, "evict + vx_upload_scan   ")
be.close()
