"""The CPU model of the GPU visibility algorithm (tests/model/wave_model.cpp: waves in reference visit
order, snapshot / resolve / trace / apply, order-scrambled merges) against the oracle, bit for bit."""
import ctypes as C
import os

import numpy as np
import pytest

from oracle.oracle import Oracle

HERE = os.path.dirname(os.path.abspath(__file__))


def run_model(grid, occ_out, endpoints, wave_cells=4096, seed=1, mode=0):
    wm = C.CDLL(os.path.join(HERE, "model", "libwave_model.so"))
    size = np.array(grid.size, np.int32)
    off = np.ascontiguousarray(grid.offset, np.float32)
    occ = np.ascontiguousarray(occ_out, np.uint8)
    ep = np.ascontiguousarray(endpoints, np.float32).reshape(-1, 3)
    o = np.zeros(grid.nvox, np.uint8)
    iv = np.zeros(grid.nvox, np.uint8)
    st = np.zeros(8, np.uint64)
    extra = np.zeros(2, np.uint64)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    wm.wm_run_mode(C.c_int(mode), p(size), C.c_float(float(grid.resolution)), p(off), p(occ), C.c_uint32(len(ep)), p(ep),
                   C.c_uint32(wave_cells), C.c_uint64(seed), p(o), p(iv), p(st), p(extra))
    if mode:
        return o, iv, st, extra
    return o, iv, st


@pytest.mark.parametrize("shape", [((0, -4, -1), (8, 4, 1), 0.5), ((0, -20, -2), (40, 20, 1), 0.6), ((-3, -2, -1), (5, 7, 2.5), 0.5),
                                   ((0, -3, 0), (4, 3, 8), 1.0)])
@pytest.mark.parametrize("dens", [0.002, 0.03, 0.25])
def test_wave_model_matches_oracle(shape, dens):
    orc = Oracle("port")
    mn, mx, res = shape
    rng = np.random.default_rng(int(dens * 1000) + len(str(shape)))
    g = orc.grid(res, mn, mx)
    n = int(g.nvox * dens) + 1
    pts = np.column_stack([rng.uniform(mn[a] - 0.5, mx[a] + 0.5, n) for a in range(3)] + [np.ones(n)]).astype(np.float32)
    eps = np.column_stack([rng.uniform(mn[a] - 2, mx[a] + 2, 6) for a in range(3)]).astype(np.float32)
    eps[0] = 0
    eps[1] = g.offset + np.float32(res) * np.array([3, 2, 1], np.float32) + np.float32(res * 0.5)
    g.insert(pts, np.ones(n, np.uint32))
    g.update_occlusions()
    for e in eps:
        g.update_invalid(e)
    occ_f, inv_f = g.flags()
    occb = (g.to_output_order(g.counts()) > 0).astype(np.uint8)
    for wave_cells in (4096, 64, 1):
        o, iv, st = run_model(g, occb, eps, wave_cells, seed=wave_cells + 3)
        assert st[7] == 0, "a visited cell had no memo"
        assert np.array_equal(o, occ_f) and np.array_equal(iv, inv_f), (wave_cells, int((o != occ_f).sum()), int((iv != inv_f).sum()))


@pytest.mark.parametrize("shape", [((0, -4, -1), (8, 4, 1), 0.5), ((0, -20, -2), (40, 20, 1), 0.6), ((-3, -2, -1), (5, 7, 2.5), 0.5)])
@pytest.mark.parametrize("dens", [0.002, 0.03, 0.25])
def test_rays_in_flight_formulation_matches_oracle(shape, dens):
    """Model mode 1 (experiments/r01_rays_in_flight.patch: the next step for the visibility kernel): the rays of an
    invalid pass stay in flight across waves, kills are applied at the end of the pass, and a pass whose rays miss
    their end voxel (endpoints ON voxel faces) is run again conservatively -- same bits as the reference."""
    orc = Oracle("port")
    mn, mx, res = shape
    rng = np.random.default_rng(int(dens * 1000) + 7 * len(str(shape)))
    g = orc.grid(res, mn, mx)
    n = int(g.nvox * dens) + 1
    pts = np.column_stack([rng.uniform(mn[a] - 0.5, mx[a] + 0.5, n) for a in range(3)] + [np.ones(n)]).astype(np.float32)
    eps = np.column_stack([rng.uniform(mn[a] - 2, mx[a] + 2, 8) for a in range(3)]).astype(np.float32)
    # endpoints exactly on voxel faces / edges / corners: rays step past the end voxel
    eps[0] = g.offset + np.float32(res) * np.array([3, 2, 1], np.float32)
    eps[1] = g.offset + np.float32(res) * np.array([2, 3.5, 1], np.float32)
    eps[2] = g.offset + np.float32(res) * np.array([1.5, 1.5, 2], np.float32)
    g.insert(pts, np.ones(n, np.uint32))
    g.update_occlusions()
    for e in eps:
        g.update_invalid(e)
    occ_f, inv_f = g.flags()
    occb = (g.to_output_order(g.counts()) > 0).astype(np.uint8)
    parked = redone = 0
    for wave_cells in (4096, 64, 1):
        o, iv, st, extra = run_model(g, occb, eps, wave_cells, seed=wave_cells + 5, mode=1)
        assert st[7] == 0
        assert np.array_equal(o, occ_f) and np.array_equal(iv, inv_f), (wave_cells, int((o != occ_f).sum()), int((iv != inv_f).sum()))
        o0, iv0, st0 = run_model(g, occb, eps, wave_cells, seed=wave_cells + 5)
        assert st[0] == st0[0] and st[1] == st0[1]      # the same rays, the same steps
        parked += int(extra[0])
        redone += int(extra[1])
    assert parked > 0
    if dens < 0.1:
        assert redone > 0, "no pass was run again: the endpoints on voxel faces no longer make rays miss their end voxel?"


@pytest.mark.parametrize("shape", [((0, -4, -1), (8, 4, 1), 0.5), ((0, -20, -2), (40, 20, 1), 0.6), ((-3, -2, -1), (5, 7, 2.5), 0.5)])
@pytest.mark.parametrize("dens", [0.002, 0.03, 0.25])
def test_mark_filter_formulation_matches_oracle(shape, dens):
    """Model mode 2 (studied in round 2, profiles/r02_memo_layout_study.txt): a ray skips the memo mark of every cell whose
    2^s x 2^s x 1 block held no live cell when the pass began -- only snapshots of LIVE cells ever read the memo, and a cell that is
    dead at the start of a pass stays dead.  Same bits as the reference, the same rays and steps, and some marks really are skipped."""
    orc = Oracle("port")
    mn, mx, res = shape
    rng = np.random.default_rng(int(dens * 1000) + 11 * len(str(shape)))
    g = orc.grid(res, mn, mx)
    n = int(g.nvox * dens) + 1
    pts = np.column_stack([rng.uniform(mn[a] - 0.5, mx[a] + 0.5, n) for a in range(3)] + [np.ones(n)]).astype(np.float32)
    eps = np.column_stack([rng.uniform(mn[a] - 2, mx[a] + 2, 6) for a in range(3)]).astype(np.float32)
    g.insert(pts, np.ones(n, np.uint32))
    g.update_occlusions()
    for e in eps:
        g.update_invalid(e)
    occ_f, inv_f = g.flags()
    occb = (g.to_output_order(g.counts()) > 0).astype(np.uint8)
    skipped = 0
    for block_shift in (1, 2, 3):
        for wave_cells in (4096, 64):
            o, iv, st, extra = run_model(g, occb, eps, wave_cells | (block_shift << 24), seed=wave_cells + block_shift, mode=2)
            assert st[7] == 0
            assert np.array_equal(o, occ_f) and np.array_equal(iv, inv_f), (block_shift, wave_cells)
            o0, iv0, st0 = run_model(g, occb, eps, wave_cells, seed=wave_cells + block_shift)
            assert st[0] == st0[0] and st[1] == st0[1] and int(extra[1]) == int(st[1])   # the same rays and steps; one mark per step
            skipped += int(extra[0])
    assert skipped > 0
