"""The kept reference host API (vxb::VoxelGrid, fillVoxelGrid, saveVoxelGrid -- voxelizer_b200/host/voxel_grid.h)
against the oracle: a gen_data.cpp-shaped caller and direct per-voxel queries.  GPU only (the facade has no CPU path)."""
import os

import numpy as np
import pytest

from oracle.oracle import Oracle
from util import cfg_path, compare_dirs, make_sequence

pytestmark = pytest.mark.gpu


def test_gen_data_written_against_the_reference_api(tmp_path):
    import voxelizer_b200 as vx
    seq = make_sequence(str(tmp_path / "seq"), 9, 16, 300, seed=11)
    cfg = cfg_path("test_small")
    n = vx.gen_data_facade(cfg, seq, str(tmp_path / "facade"))
    st = Oracle("port").gen_data(cfg, seq, str(tmp_path / "oracle"))
    assert n == st["frames"] and n > 0
    assert compare_dirs(str(tmp_path / "facade"), str(tmp_path / "oracle")) == []


def test_gen_data_facade_with_ignore_and_join(tmp_path):
    import voxelizer_b200 as vx
    seq = make_sequence(str(tmp_path / "seq"), 5, 32, 400, seed=5)
    cfg = cfg_path("settings")  # has `ignore` (and the GUI's extents)
    n = vx.gen_data_facade(cfg, seq, str(tmp_path / "facade"), max_frames=2)
    Oracle("port").gen_data(cfg, seq, str(tmp_path / "oracle"), max_frames=2)
    assert n == 2
    assert compare_dirs(str(tmp_path / "facade"), str(tmp_path / "oracle")) == []


@pytest.mark.parametrize("shape", [((0, -4, -1), (8, 4, 1), 0.5), ((-3, -2, -1), (5, 7, 2.5), 0.4)])
def test_voxel_grid_queries_match_the_reference_grid(shape):
    import voxelizer_b200 as vx
    mn, mx, res = shape
    rng = np.random.default_rng(42)
    orc = Oracle("port")
    g = orc.grid(res, mn, mx)
    n = int(g.nvox * 0.05) + 1
    pts = np.column_stack([rng.uniform(mn[a] - 0.5, mx[a] + 0.5, n) for a in range(3)] + [np.ones(n)]).astype(np.float32)
    labels = rng.integers(0, 5, n).astype(np.uint32) * 10
    eps = np.column_stack([rng.uniform(mn[a] - 1, mx[a] + 1, 4) for a in range(3)]).astype(np.float32)
    g.insert(pts, labels)
    g.update_occlusions()
    for e in eps:
        g.update_invalid(e)
    occ_o, inv_o = g.flags()                       # output order
    got = vx.facade_probe(res, mn, mx, pts, labels, eps)
    assert got["size"] == g.size and np.array_equal(got["offset"], g.offset)
    assert np.array_equal(g.to_output_order(got["occluded"]), occ_o)
    assert np.array_equal(g.to_output_order(got["invalid"]), inv_o)
    assert np.array_equal(got["free"], 1 - got["occluded"])
    assert np.array_equal(got["count"], g.counts())
    h = g.hist()                                   # (internal voxel, label, count) sorted by voxel, label
    nlab = np.bincount(h[:, 0], minlength=g.nvox)
    assert np.array_equal(got["n_labels"], nlab)
    best = np.zeros(g.nvox, np.uint32)
    lab = np.zeros(g.nvox, np.uint32)
    for v, l, c in h:                              # strict '>' in ascending label order (voxelize_utils.cpp:237-245)
        if c > best[v]:
            best[v], lab[v] = c, l
    assert np.array_equal(got["majority"], lab)
