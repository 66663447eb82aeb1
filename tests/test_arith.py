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


def test_packed_out_of_grid_test_equals_the_per_axis_rule():
    """trace_warp (vx_visibility.cu) decides "the step left the grid" with ONE test on the packed position
    P = i | j << 11 | k << 22: some field of P ^ OUT is zero, where OUT holds per axis the size (positive Step), 0
    (negative Step: Out = 0, the ray ends ON layer 0) or the wrapped -1 (negative Step from layer 0).  Mirror of that
    arithmetic against the reference's rule (VoxelGrid.cpp:286-287,307: pos == Out after the step, or pos < 0 at the
    top of the next iteration), on random walks through small and maximal grids."""
    rng = np.random.default_rng(11)
    M32 = 0xFFFFFFFF
    LOW = 1 | (1 << 11) | (1 << 22)
    HIGH = (1 << 10) | (1 << 21) | (1 << 31)

    def has_zero_field(x):
        return (((x - LOW) & M32) & (~x & M32) & HIGH) != 0

    checked = 0
    for trial in range(4000):
        n = [int(rng.choice([1, 2, 3, 7, 32, 256, 2047])), int(rng.choice([1, 2, 5, 64, 256, 2047])), int(rng.choice([1, 2, 10, 32, 64, 1023]))]
        pos = [int(rng.integers(0, n[a])) for a in range(3)]
        if rng.random() < 0.4:  # start on a boundary layer
            a = int(rng.integers(0, 3))
            pos[a] = 0 if rng.random() < 0.5 else n[a] - 1
        step = [int(rng.choice([-1, 1])) for _ in range(3)]
        out_ref = [n[a] if step[a] > 0 else 0 for a in range(3)]                      # VoxelGrid.cpp:258-259,265-266
        full = [0x7FF, 0x7FF, 0x3FF]
        out_field = [n[a] if step[a] > 0 else (full[a] if pos[a] == 0 else 0) for a in range(3)]
        OUT = out_field[0] | (out_field[1] << 11) | (out_field[2] << 22)
        P = pos[0] | (pos[1] << 11) | (pos[2] << 22)
        assert not has_zero_field(P ^ OUT)                                           # a ray that is inside is never "out"
        shift = [0, 11, 22]
        for _ in range(6000):
            a = int(rng.integers(0, 3))
            pos[a] += step[a]
            P = (P + (step[a] << shift[a])) & M32
            ref_out = pos[a] == out_ref[a] or pos[a] < 0
            assert has_zero_field(P ^ OUT) == ref_out, (n, pos, step, a)
            checked += 1
            if ref_out:
                break
            assert P == pos[0] | (pos[1] << 11) | (pos[2] << 22)                     # no field disturbed another one
    assert checked > 50000
