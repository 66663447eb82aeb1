"""Pins the oracle port (oracle/port.cpp) against the reference's own code compiled unmodified
(oracle/_ref, built from /root/reference by oracle/Makefile).  Skipped where oracle/_ref is absent."""
import os

import numpy as np
import pytest

from oracle.oracle import Oracle, have_ref
from util import cfg_path, compare_dirs, load_scan, make_sequence, random_cfg_text

pytestmark = pytest.mark.skipif(not have_ref(), reason="oracle/_ref not built (needs /root/reference)")


@pytest.fixture(scope="module")
def both():
    return Oracle("port"), Oracle("ref")


@pytest.mark.parametrize("name", ["semantic_kitti", "semantic_kitti_dense", "settings", "example", "stress_512", "test_small"])
def test_config_files(both, name):
    port, ref = both
    assert port.load_config(cfg_path(name)).as_dict() == ref.load_config(cfg_path(name)).as_dict()


def test_config_parser_corners(both, tmp_path):
    port, ref = both
    rng = np.random.default_rng(5)
    for i in range(20):
        p = tmp_path / f"c{i}.cfg"
        p.write_text(random_cfg_text(rng))
        assert port.load_config(str(p)).as_dict() == ref.load_config(str(p)).as_dict()
    # unreadable file -> defaults
    assert port.load_config(str(tmp_path / "missing.cfg")).as_dict() == ref.load_config(str(tmp_path / "missing.cfg")).as_dict()
    # bad number -> both raise
    bad = tmp_path / "bad.cfg"
    bad.write_text("voxel size: 0.2x\n")
    for o in both:
        with pytest.raises(RuntimeError):
            o.load_config(str(bad))


def test_pose_algebra_bitwise(both):
    port, ref = both
    rng = np.random.default_rng(11)
    for i in range(300):
        if i % 2:
            a = rng.uniform(-4, 4, 16).astype(np.float32)
            b = rng.uniform(-4, 4, 16).astype(np.float32)
        else:  # rigid poses far from the origin, as in a KITTI trajectory
            def rigid():
                yaw = rng.uniform(-3.2, 3.2)
                m = np.eye(4, dtype=np.float32)
                m[0, 0], m[0, 1], m[1, 0], m[1, 1] = np.cos(yaw), -np.sin(yaw), np.sin(yaw), np.cos(yaw)
                m[:3, 3] = rng.uniform(-4000, 4000, 3)
                return m.T.reshape(16).copy()
            a, b = rigid(), rigid()
        for fn in ("mat_inverse",):
            assert np.array_equal(getattr(port, fn)(a).view(np.uint32), getattr(ref, fn)(a).view(np.uint32))
        assert np.array_equal(port.mat_mul(a, b).view(np.uint32), ref.mat_mul(a, b).view(np.uint32))
        assert np.array_equal(port.endpoint(a, b).view(np.uint32), ref.endpoint(a, b).view(np.uint32))


def test_reader_poses(both, tmp_path):
    port, ref = both
    seq = make_sequence(str(tmp_path / "seq"), 9, 8, 64)
    (cp, pp), (cr, pr) = port.read_poses(seq), ref.read_poses(seq)
    assert cp == cr == 9
    assert np.array_equal(pp.view(np.uint32), pr.view(np.uint32))


def _grid_state(g):
    occ, inv = g.flags()
    return dict(counts=g.counts(), hist=g.hist(), occlusions=g.dump("occlusions"), invalid=g.dump("invalid"),
                memo=g.dump("memo"), occ=occ, inv=inv)


def test_fill_and_visibility_internals(both, tmp_path):
    """fillVoxelGrid + updateOcclusions + updateInvalid: every internal array identical."""
    port, ref = both
    seq = make_sequence(str(tmp_path / "seq"), 6, 16, 300)
    cnt, poses = ref.read_poses(seq)
    scans, labels = zip(*[load_scan(seq, t) for t in range(cnt)])
    labels = [l & 0xFFFF for l in labels]
    for cfg_name in ("test_small", "example"):
        states = []
        for o in (port, ref):
            cfg = o.load_config(cfg_path(cfg_name))
            g = o.grid(cfg.pod.voxelSize, list(cfg.pod.minExtent), list(cfg.pod.maxExtent))
            g.fill(cfg, poses[1], scans, labels, poses)
            g.update_occlusions()
            for t in range(2, cnt):
                g.update_invalid(o.endpoint(poses[1], poses[t]))
            states.append((g.size, g.offset.tobytes(), _grid_state(g)))
        assert states[0][0] == states[1][0] and states[0][1] == states[1][1]
        for k in states[0][2]:
            assert np.array_equal(states[0][2][k], states[1][2][k]), (cfg_name, k)


def test_insert_boundaries(both):
    """Points exactly on voxel faces / grid borders, NaN and inf coordinates."""
    port, ref = both
    res, mn, mx = 0.2, (0, -25.6, -2), (51.2, 25.6, 4.4)
    pts = []
    for v in (0.0, 0.2, 0.4, 0.6000000238, 51.2, 51.19999, -0.0, -1e-7, 1e-7, 25.6, -25.6, 4.4, -2.0):
        pts += [(v, 0.1, 0.1, 1), (0.1, v, 0.1, 1), (0.1, 0.1, v, 1)]
    pts += [(np.nan, 0, 0, 1), (0, np.inf, 0, 1), (-np.inf, 1, 1, 1), (3e9, 1, 1, 1), (1, -3e9, 1, 1), (1e38, 1e38, 1e38, 1)]
    pts = np.array(pts, np.float32)
    labs = np.arange(len(pts), dtype=np.uint32) % 7
    res_ = []
    for o in (port, ref):
        g = o.grid(res, mn, mx)
        g.insert(pts, labs)
        res_.append((g.counts(), g.hist()))
    assert np.array_equal(res_[0][0], res_[1][0]) and np.array_equal(res_[0][1], res_[1][1])


@pytest.mark.parametrize("shape", [((0, -4, -1), (8, 4, 1), 0.5), ((-3, -2, -1), (5, 7, 2.5), 0.5), ((0, -3, 0), (4, 3, 8), 1.0),
                                   ((0, -20, -2), (40, 20, 1), 0.3)])
def test_visibility_random_grids(both, shape):
    port, ref = both
    mn, mx, res = shape
    rng = np.random.default_rng(42)
    for dens in (0.003, 0.05, 0.3):
        g0 = port.grid(res, mn, mx)
        n = int(g0.nvox * dens) + 1
        pts = np.column_stack([rng.uniform(mn[a] - 0.5, mx[a] + 0.5, n) for a in range(3)] + [np.ones(n)]).astype(np.float32)
        eps = np.column_stack([rng.uniform(mn[a] - 2, mx[a] + 2, 5) for a in range(3)]).astype(np.float32)
        eps[0] = 0
        eps[1] = g0.offset + np.float32(res) * np.array([3, 2, 1], np.float32) + np.float32(res * 0.5)  # a voxel centre
        out = []
        for o in (port, ref):
            g = o.grid(res, mn, mx)
            g.insert(pts, np.ones(n, np.uint32))
            g.update_occlusions()
            for e in eps:
                g.update_invalid(e)
            out.append((g.dump("occlusions"), g.dump("invalid"), g.dump("memo")))
        for a, b in zip(*out):
            assert np.array_equal(a, b)


@pytest.mark.parametrize("cfg,n,beams,az", [("test_small", 14, 16, 300), ("example", 8, 24, 400), ("settings", 4, 16, 300)])
def test_gen_data_whole_program(both, tmp_path, cfg, n, beams, az):
    port, ref = both
    seq = make_sequence(str(tmp_path / "seq"), n, beams, az)
    port.gen_data(cfg_path(cfg), seq, str(tmp_path / "port"))
    ref.gen_data(cfg_path(cfg), seq, str(tmp_path / "ref"))
    assert compare_dirs(str(tmp_path / "port"), str(tmp_path / "ref")) == []
    assert len(os.listdir(tmp_path / "ref")) > 0


def test_gen_data_without_labels_skips_everything(both, tmp_path):
    """Missing labels/ is created zero-filled (KittiReader.cpp:31-50) and every frame fails the gate."""
    port, ref = both
    for o, name in ((port, "p"), (ref, "r")):
        seq = make_sequence(str(tmp_path / f"seq_{name}"), 5, 8, 100, labels=False)
        o.gen_data(cfg_path("test_small"), seq, str(tmp_path / f"out_{name}"))
        assert os.listdir(tmp_path / f"out_{name}") == []
        labs = sorted(os.listdir(os.path.join(seq, "labels")))
        assert len(labs) == 5
        assert os.path.getsize(os.path.join(seq, "labels", labs[0])) == os.path.getsize(os.path.join(seq, "velodyne", "000000.bin")) // 4
