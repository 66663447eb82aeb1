import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
if os.path.join(ROOT, "tests") not in sys.path:
    sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def _native_libraries():
    """Everything is built in-tree; (re)build what is stale.  The prebuilt files travel to the GPU box."""
    from voxelizer_b200 import build as vxbuild
    from oracle import oracle as orc

    vxbuild.build_all()
    orc.build()  # the port always; oracle/_ref only where /root/reference exists
    import subprocess
    model = os.path.join(ROOT, "tests", "model", "libwave_model.so")
    src = os.path.join(ROOT, "tests", "model", "wave_model.cpp")
    hdr = os.path.join(ROOT, "voxelizer_b200", "csrc", "vx_dda.h")
    if not os.path.exists(model) or os.path.getmtime(model) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.check_call(["g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-ffp-contract=off", src, "-o", model])
    from util import build_callers
    build_callers()  # tests/callers: reference-shaped callers of the kept host API (links libvxhost.so)
    yield
