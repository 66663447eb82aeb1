"""CPU checks of the two arithmetic shortcuts the CUDA path takes (restated in numpy with the same operation
order; fp32 ops in numpy are individually rounded, like the kernel's __f*_rn intrinsics):

* vx_dda.h computes the reference's two double-precision intermediates of the ray set-up
  (VoxelGrid.h:93-97, VoxelGrid.cpp:249,257,264) in fp32;
* vx_fill.cu `floor_div` replaces floorf(IEEE v / res) (VoxelGrid.cpp:47-49) by a multiply whenever the floor
  is provably the same.
"""
import numpy as np
import pytest

F = np.float32


@pytest.mark.parametrize("res", [0.2, 0.1, 0.3, 0.5, 0.123456])
def test_ray_setup_in_fp32_equals_the_double_intermediates(res):
    rng = np.random.default_rng(int(res * 1e6))
    res = F(res)
    n = 2_000_000
    d = (rng.standard_normal(n) * rng.choice([1e-3, 1.0, 50.0, 1e4], n)).astype(F)
    d = d[d != 0]
    half64 = 0.5 * np.float64(res)                                     # double halfResolution = 0.5 * resolution_
    ref_tmax = (half64 / np.abs(d).astype(np.float64)).astype(F)       # (+-half / dir) in double, narrowed
    tdelta = (res / np.abs(d)).astype(F)                               # fp32 DeltaT
    assert np.array_equal(ref_tmax, (F(0.5) * tdelta).astype(F))       # what vx_axis_setup does
    off = F(-25.6)
    i = rng.integers(0, 600, len(d)).astype(np.int32)
    b = (off + (i.astype(F) * res).astype(F)).astype(F)
    ref_centre = (b.astype(np.float64) + half64).astype(F)             # voxel2position
    assert np.array_equal(ref_centre, (b + (F(0.5) * res)).astype(F))  # vx_voxel_centre


def floor_div(v, res):
    rcp = F(1.0) / res
    q0 = (v * rcp).astype(F)
    n = np.floor(q0)
    d = (q0 - n).astype(F)
    with np.errstate(invalid="ignore"):
        fast = np.abs((d - F(0.5)).astype(F)) < (F(0.5) - (np.abs(q0) * F(4.76837158e-7)).astype(F)).astype(F)  # floor_div3
    exact = np.floor((v / res).astype(F))
    return np.where(fast, n, exact), fast, exact


@pytest.mark.parametrize("res", [0.2, 0.1, 0.3, 0.5, 0.05, 0.123456, 1.0])
def test_division_free_floor_matches_ieee_division(res):
    rng = np.random.default_rng(int(res * 1e6) + 7)
    res = F(res)
    n = 2_000_000
    k = rng.integers(-400, 400, n).astype(np.float64)
    near = (k * np.float64(res)).astype(F)
    samples = [
        rng.uniform(-80, 80, n).astype(F),                              # ordinary coordinates
        (k * np.float64(res) + rng.normal(0, 1e-6, n)).astype(F),        # on and around voxel faces
        np.nextafter(near, F(np.inf)), np.nextafter(near, F(-np.inf)), near,
        (rng.standard_normal(n) * 1e5).astype(F),                        # far outside the grid
        np.array([0.0, -0.0, np.inf, -np.inf, np.nan, 1e38, -1e38, 1e-45], F),
    ]
    fast_share = []
    for v in samples:
        got, fast, exact = floor_div(v, res)
        ok = (got == exact) | (np.isnan(got) & np.isnan(exact))
        assert ok.all()
        assert not fast[~np.isfinite(v)].any()                          # NaN / infinity always take the IEEE path
        fast_share.append(fast.mean())
    assert fast_share[0] > 0.999                                         # the fallback is rare for ordinary points
