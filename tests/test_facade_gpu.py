"""The kept reference host API (vxb::VoxelGrid, fillVoxelGrid, saveVoxelGrid -- voxelizer_b200/host/voxel_grid.h)
against the oracle, driven by callers shaped like the reference's own two users of the path (tests/callers/callers.cpp):
the gen_data frame loop and Mainframe::buildVoxelGrids + the voxel lists the GUI draws; plus direct per-voxel
queries and the public VoxelGrid::occludedBy.  GPU only (the facade has no CPU path)."""
import numpy as np
import pytest

import voxelizer_b200 as vx
from frames import oracle_frame
from oracle.oracle import Oracle, have_ref
from util import build_voxel_grids, cfg_path, compare_dirs, gen_data_frame_by_frame, load_scan, make_sequence, voxel_grid_probe

pytestmark = pytest.mark.gpu

KINDS = ["port", "ref"] if have_ref() else ["port"]


@pytest.fixture(scope="module", params=KINDS)
def orc(request):
    return Oracle(request.param)


def test_gen_data_written_against_the_reference_api(tmp_path):
    seq = make_sequence(str(tmp_path / "seq"), 9, 16, 300, seed=11)
    cfg = cfg_path("test_small")
    n = gen_data_frame_by_frame(cfg, seq, str(tmp_path / "facade"))
    st = Oracle("port").gen_data(cfg, seq, str(tmp_path / "oracle"))
    assert n == st["frames"] and n > 0
    assert compare_dirs(str(tmp_path / "facade"), str(tmp_path / "oracle")) == []


def test_gen_data_facade_with_ignore_and_join(tmp_path):
    seq = make_sequence(str(tmp_path / "seq"), 5, 32, 400, seed=5)
    cfg = cfg_path("settings")  # has `ignore` (and the GUI's extents)
    n = gen_data_frame_by_frame(cfg, seq, str(tmp_path / "facade"), max_frames=2)
    Oracle("port").gen_data(cfg, seq, str(tmp_path / "oracle"), max_frames=2)
    assert n == 2
    assert compare_dirs(str(tmp_path / "facade"), str(tmp_path / "oracle")) == []


@pytest.mark.parametrize("cfg_name,n,beams,az,target", [("settings", 6, 32, 500, 2), ("semantic_kitti", 12, 64, 1500, 3)])
def test_build_voxel_grids_like_the_gui(orc, tmp_path, cfg_name, n, beams, az, target):
    """Mainframe::buildVoxelGrids (Mainframe.cpp:279-332) + extractLabeledVoxels / updateOccludedVoxels /
    updateInvalidVoxels (:369-447): the lists the GUI would draw, and a latency budget for the rebuild."""
    seq = make_sequence(str(tmp_path / "seq"), n, beams, az, seed=21)
    cnt, poses = vx.read_poses(seq)
    scans, raw = zip(*[load_scan(seq, t) for t in range(cnt)])
    labels = [l & 0xFFFF for l in raw]
    cfg = vx.Config(cfg_path(cfg_name))
    pb, pe, qb, qe = vx.window(cnt, cfg.pod.priorScans, cfg.pod.pastScans, target)
    o = oracle_frame(orc, cfg_path(cfg_name), list(scans), labels, poses, list(range(pb, pe)), list(range(qb, qe)))
    sx, sy, sz = o["size"]
    got = build_voxel_grids(cfg_path(cfg_name), seq, target, sx * sy * sz)

    def expect_labeled(hist):  # voxels with count > 0 in x, y, z order with the majority label (strict '>' over ascending labels)
        L = np.unique(hist[:, 0])
        best = {}
        for v, l, c in hist:
            if v not in best or c > best[v][0]:
                best[v] = (c, l)
        rows = [(v // (sy * sz), (v // sz) % sy, v % sz, best[v][1]) for v in L]
        return np.array(rows, np.uint32).reshape(-1, 4)

    assert np.array_equal(got["labeled_past"], expect_labeled(o["hist_target"]))
    assert np.array_equal(got["labeled_prior"], expect_labeled(o["hist_input"]))
    assert np.array_equal(got["occluded_past"], np.flatnonzero(o["occluded"]).astype(np.uint32))
    assert np.array_equal(got["invalid_past"], np.flatnonzero(o["invalid"]).astype(np.uint32))
    assert len(got["occluded_prior"]) > 0
    assert got["seconds"] < 20.0, got["seconds"]   # interactive rebuild budget (context creation included)
    print(f"buildVoxelGrids({cfg_name}, {qe - qb} past scans): {got['seconds']:.2f} s")


@pytest.mark.parametrize("shape", [((0, -4, -1), (8, 4, 1), 0.5), ((-3, -2, -1), (5, 7, 2.5), 0.4)])
def test_voxel_grid_queries_and_public_rays_match_the_reference_grid(orc, shape):
    mn, mx, res = shape
    rng = np.random.default_rng(42)
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
    sx, sy, sz = g.size
    # public occludedBy: random starts (some outside the grid) towards random endpoints, the origin, far away, a voxel face
    rays = []
    for _ in range(60):
        start = tuple(int(rng.integers(-1, s + 1)) for s in (sx, sy, sz))
        end = tuple(float(rng.uniform(mn[a] - 2, mx[a] + 2)) for a in range(3))
        rays.append((start, end))
    rays += [((sx - 1, 0, sz - 1), (0.0, 0.0, 0.0)), ((sx // 2, sy // 2, 0), (1e4, -1e4, 3.0)),
             ((1, 1, 1), (float(g.offset[0] + 3 * res), float(g.offset[1] + 2 * res), float(g.offset[2] + res)))]
    got = voxel_grid_probe(res, mn, mx, pts, labels, eps, g.nvox, rays=rays)
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
    hits = 0
    for (start, end), (occluder, visited) in zip(rays, got["rays"]):
        o_occ, o_vis = g.occluded_by(*start, end)
        assert occluder == o_occ, (start, end)
        assert np.array_equal(visited, o_vis), (start, end)
        hits += occluder >= 0
    assert hits > 0


@pytest.mark.parametrize("n_endpoints", [0, 3])
@pytest.mark.parametrize("shape", [((0, -4, -1), (8, 4, 3), 0.5), ((0, -10, -2), (20, 10, 4.4), 0.4)])
def test_insert_occlusion_labels_matches_the_reference_grid(orc, tmp_path, shape, n_endpoints):
    """VoxelGrid::insertOcclusionLabels (VoxelGrid.cpp:75-96) in the order of the reference's commented-out call sites
    (gen_data.cpp:102-116): fill, updateOcclusions, insertOcclusionLabels, updateInvalid x n; every per-voxel query and the
    four saveVoxelGrid images.  Filled voxels are occupied for the invalid passes and keep invalid_ = min(occluder, own index)."""
    mn, mx, res = shape
    rng = np.random.default_rng(7 + n_endpoints)
    g = orc.grid(res, mn, mx)
    sx, sy, sz = g.size
    # structures with empty space below them: blobs in the upper half of the grid, plus a few ground points
    n = int(g.nvox * 0.03) + 1
    pts = np.column_stack([rng.uniform(mn[0] + 1, mx[0], n), rng.uniform(mn[1], mx[1], n),
                           rng.uniform(mn[2] + 0.5 * (mx[2] - mn[2]), mx[2], n), np.ones(n)]).astype(np.float32)
    ground = np.column_stack([rng.uniform(mn[0], mx[0], n // 3), rng.uniform(mn[1], mx[1], n // 3),
                              np.full(n // 3, mn[2] + 0.1), np.ones(n // 3)]).astype(np.float32)
    pts = np.vstack([pts, ground])
    labels = (rng.integers(1, 6, len(pts)) * 10).astype(np.uint32)
    eps = np.column_stack([rng.uniform(mn[a], mx[a] + 2, n_endpoints) for a in range(3)]).astype(np.float32).reshape(-1, 3)
    g.insert(pts, labels)
    g.update_occlusions()
    before = int((g.counts() > 0).sum())
    g.insert_occlusion_labels()
    assert int((g.counts() > 0).sum()) > before + 20          # the case does fill voxels
    for e in eps:
        g.update_invalid(e)
    occ_o, inv_o = g.flags()
    got = voxel_grid_probe(res, mn, mx, pts, labels, eps, g.nvox, insert_occlusion_labels=True)
    assert np.array_equal(got["count"], g.counts())
    h = g.hist()
    assert np.array_equal(got["n_labels"], np.bincount(h[:, 0], minlength=g.nvox))
    best = np.zeros(g.nvox, np.uint32)
    lab = np.zeros(g.nvox, np.uint32)
    for v, l, c in h:
        if c > best[v]:
            best[v], lab[v] = c, l
    assert np.array_equal(got["majority"], lab)
    assert np.array_equal(g.to_output_order(got["occluded"]), occ_o)
    assert np.array_equal(g.to_output_order(got["invalid"]), inv_o)
    assert inv_o.sum() > 0
    # the files saveVoxelGrid writes for this grid in both modes
    g.save(str(tmp_path), "f", "input")
    g.save(str(tmp_path), "f", "target")
    img = got["images"]
    assert np.array_equal(img["label"], np.fromfile(str(tmp_path / "f.label"), np.uint16))
    assert np.array_equal(img["bin"], np.fromfile(str(tmp_path / "f.bin"), np.uint8))
    assert np.array_equal(img["occluded"], np.fromfile(str(tmp_path / "f.occluded"), np.uint8))
    assert np.array_equal(img["invalid"], np.fromfile(str(tmp_path / "f.invalid"), np.uint8))
