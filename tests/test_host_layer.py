"""The C++ host layer of the product (voxelizer_b200/host: Config, KittiReader, pose algebra, target
planning) against the oracle port -- CPU only."""
import numpy as np
import pytest

import voxelizer_b200 as vx
from oracle.oracle import Oracle
from util import cfg_path, make_sequence, random_cfg_text


@pytest.fixture(scope="module")
def port():
    return Oracle("port")


@pytest.mark.parametrize("name", ["semantic_kitti", "semantic_kitti_dense", "settings", "example", "stress_512", "test_small"])
def test_parse_configuration(port, name):
    assert vx.Config(cfg_path(name)).as_dict() == port.load_config(cfg_path(name)).as_dict()


def test_parse_configuration_corners(port, tmp_path):
    rng = np.random.default_rng(17)
    for i in range(20):
        p = tmp_path / f"c{i}.cfg"
        p.write_text(random_cfg_text(rng))
        assert vx.Config(str(p)).as_dict() == port.load_config(str(p)).as_dict()
    assert vx.Config(str(tmp_path / "nope.cfg")).as_dict() == port.load_config(str(tmp_path / "nope.cfg")).as_dict()
    # empty lists: boost::tokenizer yields no token for an empty range, so both keys are no-ops in the reference
    empty = tmp_path / "empty_lists.cfg"
    empty.write_text("voxel size: 0.5\nignore: []\njoin: []\npast scans: 3\n")
    d = vx.Config(str(empty)).as_dict()
    assert d == port.load_config(str(empty)).as_dict()
    assert d["filtered"] == [] and d["join_pairs"] == [] and d["pastScans"] == 3
    bad = tmp_path / "bad.cfg"
    bad.write_text("max range: 7 0\n")
    with pytest.raises(vx.VxError):
        vx.Config(str(bad))


def test_pose_algebra(port):
    rng = np.random.default_rng(3)
    for i in range(300):
        a = rng.uniform(-5, 5, 16).astype(np.float32)
        b = rng.uniform(-5, 5, 16).astype(np.float32)
        if i % 3 == 0:
            a[[3, 7, 11]] = 0
            a[15] = 1
            a[12:15] = rng.uniform(-4000, 4000, 3)
        assert np.array_equal(vx.mat_inverse(a).view(np.uint32), port.mat_inverse(a).view(np.uint32))
        assert np.array_equal(vx.mat_mul(a, b).view(np.uint32), port.mat_mul(a, b).view(np.uint32))
        ap, ep = vx.frame_transforms(a, np.stack([b, a]))
        assert np.array_equal(ap[0].view(np.uint32), port.mat_mul(port.mat_inverse(a), b).view(np.uint32))
        assert np.array_equal(ep[0].view(np.uint32), port.endpoint(a, b).view(np.uint32))


def test_reader_and_targets(port, tmp_path):
    seq = make_sequence(str(tmp_path / "seq"), 23, 8, 64)
    cnt, poses = vx.read_poses(seq)
    cnt_o, poses_o = port.read_poses(seq)
    assert cnt == cnt_o == 23
    assert np.array_equal(poses.view(np.uint32), poses_o.view(np.uint32))
    # the target list is what the oracle's gen_data walks: compare through the files it writes
    import os
    for cfg_name in ("test_small", "semantic_kitti"):
        cfg = vx.Config(cfg_path(cfg_name))
        targets = vx.plan_targets(cfg, cnt, poses)
        out = tmp_path / f"out_{cfg_name}"
        port.gen_data(cfg_path(cfg_name), seq, str(out))
        written = sorted({int(f.split(".")[0]) for f in os.listdir(out)})
        assert written == targets.tolist()


def test_direct_loader_matches_reference_shaped_loader(tmp_path):
    """KittiReader::loadInto (what the pipeline's reader threads call) against KittiReader::load."""
    import os
    seq = make_sequence(str(tmp_path / "seq"), 4, 8, 64)
    for t in range(4):
        a = vx.read_scan(seq, t, direct=False)
        b = vx.read_scan(seq, t, direct=True)
        xyzr, raw = np.fromfile(os.path.join(seq, "velodyne", f"{t:06d}.bin"), np.float32).reshape(-1, 4), \
            np.fromfile(os.path.join(seq, "labels", f"{t:06d}.label"), np.uint32)
        assert np.array_equal(a[0].view(np.uint32), xyzr.view(np.uint32)) and np.array_equal(b[0].view(np.uint32), xyzr.view(np.uint32))
        assert np.array_equal(a[1], raw & 0xFFFF) and np.array_equal(b[1], raw)   # the device applies the mask to raw labels
        assert a[2] == b[2] == int(((raw & 0xFFFF) > 0).sum())
        assert (raw >> 16).any()                                                   # the generator does write instance ids
    assert vx.read_scan(seq, 0, direct=True, capacity=10) is None                 # does not fit: the pipeline falls back
    with open(os.path.join(seq, "labels", "000002.label"), "ab") as f:           # one label too many
        f.write(b"\0\0\0\0")
    for direct in (False, True):
        with pytest.raises(vx.VxError, match="Inconsistent number of labels"):
            vx.read_scan(seq, 2, direct=direct)


def test_windows_match_reader_semantics():
    # last scan: the past window is {N-1} (the last scan is inserted twice), SURVEY.md A.5
    assert vx.window(10, 0, 70, 9) == (9, 10, 9, 10)
    assert vx.window(10, 2, 3, 0) == (0, 1, 1, 4)
    assert vx.window(10, 2, 3, 5) == (3, 6, 6, 9)
    assert vx.window(10, 2, 3, 8) == (6, 9, 9, 10)


def test_missing_inputs_raise(tmp_path):
    with pytest.raises(vx.VxError, match="Missing velodyne"):
        vx.read_poses(str(tmp_path))
    seq = make_sequence(str(tmp_path / "seq"), 3, 8, 32)
    import os
    os.remove(os.path.join(seq, "calib.txt"))
    with pytest.raises(vx.VxError, match="Missing calibration file"):
        vx.read_poses(seq)
