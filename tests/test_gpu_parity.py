"""GPU parity tests proper: the CUDA path, called through the C ABI, against the oracle on the same
seeded inputs -- bit-exact (integer / byte / index work; the fp32 voxel index must match exactly too)."""
import os

import numpy as np
import pytest

import voxelizer_b200 as vx
from frames import assert_frames_equal, oracle_frame, product_frame
from oracle.oracle import Oracle, have_ref
from util import cfg_path, compare_dirs, load_scan, make_sequence

pytestmark = pytest.mark.gpu


# every parity test runs against the builder's restatement AND, where it travelled with the snapshot, against the
# reference's own translation units (oracle/_ref)
KINDS = ["port", "ref"] if have_ref() else ["port"]


@pytest.fixture(scope="module", params=KINDS)
def orc(request):
    return Oracle(request.param)


@pytest.fixture(params=[1, 2, 3], ids=["throughput-shape", "latency-shape", "middle-shape"])
def vis_mode(request):
    """The three shapes of the traversal kernel (vx_options.vis_mode): 7 x 128-thread CTAs per SM with waves of <= 1280 live cells,
    one 512-thread CTA per SM with whole-face waves, and 4 x 256-thread CTAs per SM with waves of <= 2560 live cells.  Left to
    itself the library picks by batch size."""
    return request.param


def _load_sequence(seq):
    cnt, poses = vx.read_poses(seq)
    scans, labels = zip(*[load_scan(seq, t) for t in range(cnt)])
    return cnt, poses, list(scans), [l & 0xFFFF for l in labels], list(labels)


def _backend(cfg_name, scans, raw_labels, **kw):
    be = vx.Backend.from_config(vx.Config(cfg_path(cfg_name)), debug=True, **kw)
    for t, (s, l) in enumerate(zip(scans, raw_labels)):
        be.upload_scan(t, s, l)  # raw labels: the library applies & 0xFFFF like the reader
    return be


@pytest.mark.parametrize("cfg_name,n,beams,az", [("test_small", 10, 16, 300), ("example", 8, 32, 500), ("settings", 5, 32, 600),
                                                  ("semantic_kitti", 7, 64, 1200)])
def test_frame_internals(orc, tmp_path, cfg_name, n, beams, az, vis_mode):
    """histograms, majority labels, occupancy, occlusions_, invalid_ and the packed outputs of single frames."""
    seq = make_sequence(str(tmp_path / "seq"), n, beams, az)
    cnt, poses, scans, labels, raw = _load_sequence(seq)
    cfg = vx.Config(cfg_path(cfg_name))
    be = _backend(cfg_name, scans, raw, vis_mode=vis_mode)
    for target in (0, cnt // 2, cnt - 1):  # cnt-1: the last scan is its own past window
        pb, pe, qb, qe = vx.window(cnt, cfg.pod.priorScans, cfg.pod.pastScans, target)
        prior, past = list(range(pb, pe)), list(range(qb, qe))
        o = oracle_frame(orc, cfg_path(cfg_name), scans, labels, poses, prior, past)
        p = product_frame(be, poses, prior, past)
        assert be.size == tuple(o["size"]) and np.array_equal(be.offset.view(np.uint32), np.asarray(o["offset"], np.float32).view(np.uint32))
        assert np.array_equal(p["endpoints"].view(np.uint32), o["endpoints"].view(np.uint32))
        assert_frames_equal(p, o)
    be.close()


def test_adversarial_points(orc, vis_mode):
    """Boundary ranges, voxel faces, NaN / inf, label ties, majority label 0, instance bits, empty scans."""
    cfg_name = "test_small"  # min range 2.5, max range 40, ignore [0, 30], join {0:[1]}, {10:[252,13]}; res 0.8
    rng = np.random.default_rng(1)
    pts = []
    # exactly at the range limits (kept: the reference rejects only range < min or > max)
    for r in (2.5, np.nextafter(np.float32(2.5), 0), np.nextafter(np.float32(2.5), 9), 40.0, np.nextafter(np.float32(40), 99)):
        pts += [(0, float(r), 0, 0.5), (float(r) * 0.6, float(r) * 0.8, 0, 0.5)]
    # ego-vehicle box borders
    pts += [(3.0, 0.5, 1, 0), (2.9999, 1.9999, 1, 0), (-2.0, 0.0, 3.5, 0), (-1.99, 2.0, 3.5, 0), (2.5, -2.0, 4.0, 0), (4.0, 0.1, 0, 0)]
    # voxel faces of the 0.8 m grid (offset (0,-12.8,-2)) and grid borders
    for v in (0.8, 1.6, 2.4000000954, 3.2, 25.6, 25.599998, 12.8, -12.8, 4.4, -2.0, 0.0, -0.0):
        pts += [(v, 3.0, 0.3, 0), (5.0, v, 0.3, 0), (5.0, 3.0, v, 0)]
    pts += [(np.nan, 5, 0, 0), (5, np.inf, 0, 0), (-np.inf, 5, 0, 0), (3e9, 5, 5, 0), (1e38, 1e38, 1e38, 0), (5, 5, np.nan, 0)]
    pts = np.array(pts, np.float32)
    lab = (rng.integers(0, 6, len(pts)) * 10).astype(np.uint32)
    # ties / majority label 0 / ignored + joined labels / instance bits in one voxel cluster at (6.0..6.7, 4.0..4.7, 0.5)
    cluster = [((6.1, 4.1, 0.5, 0), 40), ((6.2, 4.2, 0.5, 0), 50), ((6.3, 4.3, 0.5, 0), 50), ((6.4, 4.4, 0.5, 0), 40),   # tie 40/50 -> 40
               ((7.0, 4.1, 0.5, 0), 1), ((7.1, 4.2, 0.5, 0), 13), ((7.2, 4.3, 0.5, 0), 252), ((7.3, 4.2, 0.5, 0), 30),  # 1->0 ignored, 13/252->10 kept as 13/252, 30 ignored
               ((8.1, 4.1, 0.5, 0), 2), ((8.2, 4.1, 0.5, 0), 2), ((8.3, 4.1, 0.5, 0), 0x70002), ((8.4, 4.1, 0.5, 0), 0xFFFF0030)]  # instance bits
    pts = np.vstack([pts, np.array([c[0] for c in cluster], np.float32)])
    lab = np.concatenate([lab, np.array([c[1] for c in cluster], np.uint32)])
    scans = [pts, np.zeros((0, 4), np.float32), pts[::-1].copy() + np.float32([0.37, -0.21, 0.05, 0])]
    raw = [lab, np.zeros(0, np.uint32), lab[::-1].copy()]
    labels = [l & 0xFFFF for l in raw]
    ident = np.eye(4, dtype=np.float32).reshape(16)
    shift = np.eye(4, dtype=np.float32)
    shift[:3, 3] = (1.25, -0.5, 0.1)
    poses = np.stack([ident, shift.T.reshape(16), ident])
    be = _backend(cfg_name, scans, raw, vis_mode=vis_mode)
    for prior, past in (([0], [1, 2]), ([0, 1], [2]), ([1], [1]), ([2], [])):
        o = oracle_frame(orc, cfg_path(cfg_name), scans, labels, poses, prior, past)
        p = product_frame(be, poses, prior, past)
        assert_frames_equal(p, o)
    be.close()


def test_endpoints_everywhere(orc, tmp_path, vis_mode):
    """updateInvalid with endpoints outside the grid, on voxel centres / faces / corners and at the origin."""
    seq = make_sequence(str(tmp_path / "seq"), 3, 32, 500)
    cnt, poses, scans, labels, raw = _load_sequence(seq)
    for cfg_name in ("example", "test_small"):
        be = _backend(cfg_name, scans, raw, vis_mode=vis_mode)
        res, off = float(be.resolution), be.offset.astype(np.float64)
        sx, sy, sz = be.size
        eps = [(0, 0, 0), (70, 3, 1), (-5, -30, 9), (1e4, 1e4, 1e4), tuple(off + res * np.array([3.5, 2.5, 1.5])), tuple(off + res * np.array([4, 4, 2])),
               tuple(off), tuple(off + res * np.array([sx, sy, sz])), tuple(off + res * np.array([sx - 0.5, 0.5, sz - 0.5])), (12.3, -4.56, 0.789),
               (20, 0, -1.9999), (0.0001, 0.0001, 0.0001)]
        eps = np.array(eps, np.float32)
        o = oracle_frame(orc, cfg_path(cfg_name), scans, labels, poses, [0], [1, 2], extra_endpoints=eps)
        p = product_frame(be, poses, [0], [1, 2], extra_endpoints=eps)
        assert_frames_equal(p, o)
        be.close()


def test_random_occupancy_visibility(orc, vis_mode):
    """Dense random occupancy on odd-sized grids: stresses in-wave resolution and the multi-word columns."""
    rng = np.random.default_rng(9)
    for (mn, mx, res), dens in ((((0, -4, -1), (8, 4, 1), 0.5), 0.2), (((-3, -2, -1), (5, 7, 2.5), 0.5), 0.05), (((0, -3, 0), (4, 3, 8), 1.0), 0.3),
                                (((0, -10, -2), (20, 10, 4.6), 0.3), 0.02)):
        g = orc.grid(res, mn, mx)
        n = int(g.nvox * dens) + 1
        pts = np.column_stack([rng.uniform(mn[a], mx[a], n) for a in range(3)] + [np.zeros(n)]).astype(np.float32)
        lab = rng.integers(1, 5, n).astype(np.uint32)
        eps = np.column_stack([rng.uniform(mn[a] - 2, mx[a] + 2, 7) for a in range(3)]).astype(np.float32)
        be = vx.Backend(res, mn, mx, 0.0, 1e9, hidecar=False, debug=True, vis_mode=vis_mode)
        be.upload_scan(0, pts, lab)
        ident = np.eye(4, dtype=np.float32).reshape(1, 16)
        spec = vx.FrameSpec(np.array([0], np.uint32), ident, np.zeros(0, np.uint32), np.zeros((0, 16), np.float32), eps)
        out = be.process([spec])[0]
        g.insert(np.column_stack([pts[:, :3], np.ones(n, np.float32)]), lab)
        g.update_occlusions()
        for e in eps:
            g.update_invalid(e)
        occl, inv = g.flags()
        assert np.array_equal(be.dump(0, vx.DUMP_OCCLUDED), occl)
        assert np.array_equal(be.dump(0, vx.DUMP_INVALID), inv)
        be.close()


@pytest.mark.parametrize("cfg_name,n,beams,az", [("test_small", 17, 16, 300), ("example", 12, 32, 600), ("semantic_kitti", 11, 64, 2000),
                                                  ("semantic_kitti_dense", 4, 64, 1000)])
def test_gen_data_whole_program(orc, tmp_path, cfg_name, n, beams, az):
    """vx_gen_data's files == the oracle's gen_data files (file list and bytes)."""
    seq = make_sequence(str(tmp_path / "seq"), n, beams, az)
    orc.gen_data(cfg_path(cfg_name), seq, str(tmp_path / "oracle"))  # kind "ref": the reference BINARY (oracle/_ref/ref_gen_data)
    st = vx.gen_data(cfg_path(cfg_name), seq, str(tmp_path / "gpu"))
    assert compare_dirs(str(tmp_path / "gpu"), str(tmp_path / "oracle")) == []
    assert st["frames"] * 4 == len(os.listdir(tmp_path / "oracle"))


def test_gen_data_executable(orc, tmp_path):
    """The drop-in executable, with the reference's positional arguments."""
    import subprocess
    seq = make_sequence(str(tmp_path / "seq"), 9, 16, 300)
    orc.gen_data(cfg_path("test_small"), seq, str(tmp_path / "oracle"))
    subprocess.check_call([vx.GEN_DATA_EXE, cfg_path("test_small"), seq, str(tmp_path / "gpu"), "--quiet"])
    assert compare_dirs(str(tmp_path / "gpu"), str(tmp_path / "oracle")) == []


def test_size_independent_properties(tmp_path):
    """Full-size grid (256x256x32): batch size, sharding and repetition do not change a byte; set relations
    between the four products hold."""
    seq = make_sequence(str(tmp_path / "seq"), 16, 64, 2000)
    cfg = cfg_path("semantic_kitti_dense")
    a = vx.gen_data(cfg, seq, str(tmp_path / "a"))
    vx.gen_data(cfg, seq, str(tmp_path / "b"), frames_in_flight=3)
    vx.gen_data(cfg, seq, str(tmp_path / "c"), shard=(0, 2))
    vx.gen_data(cfg, seq, str(tmp_path / "c"), shard=(1, 2))
    assert a["frames"] == 16
    assert compare_dirs(str(tmp_path / "a"), str(tmp_path / "b")) == []
    assert compare_dirs(str(tmp_path / "a"), str(tmp_path / "c")) == []
    for t in (0, 7, 15):
        rd = lambda ext, dt: np.fromfile(str(tmp_path / "a" / f"{t:06d}.{ext}"), dt)
        label = rd("label", np.uint16)
        unpack = lambda ext: np.unpackbits(rd(ext, np.uint8), bitorder="big").astype(bool)
        b, occl, inv = unpack("bin"), unpack("occluded"), unpack("invalid")
        assert label.size == 256 * 256 * 32 and b.size == label.size
        assert not (inv & (label > 0)).any()      # labelled voxels are occupied, occupied voxels are never invalid
        assert not (inv & ~occl).any()            # invalid implies occluded from the origin
        assert ((label > 0) <= occl).all()        # occupied voxels count as occluded
        assert (b <= occl).all()                  # the target scan's own voxels are occupied in the target grid


def test_table_growth_is_transparent(orc, tmp_path):
    """A histogram table that is far too small overflows, is grown and the frame redone -- same bytes."""
    seq = make_sequence(str(tmp_path / "seq"), 5, 32, 600)
    cnt, poses, scans, labels, raw = _load_sequence(seq)
    be = _backend("example", scans, raw, table_log2_capacity=10)
    o = oracle_frame(orc, cfg_path("example"), scans, labels, poses, [0], [1, 2, 3, 4])
    p = product_frame(be, poses, [0], [1, 2, 3, 4])
    assert_frames_equal(p, o, what=("label", "bin", "occupancy", "occluded", "invalid", "occluded_packed", "invalid_packed"))
    be.close()


def test_unknown_scan_is_an_error():
    be = vx.Backend(0.5, (0, -4, -1), (8, 4, 1), 2.5, 25.0)
    ident = np.eye(4, dtype=np.float32).reshape(1, 16)
    spec = vx.FrameSpec(np.array([7], np.uint32), ident, np.zeros(0, np.uint32), np.zeros((0, 16), np.float32), np.zeros((0, 3), np.float32))
    with pytest.raises(vx.VxError, match="not resident"):
        be.process([spec])
    be.close()


def test_stress_grid_512(orc, tmp_path, vis_mode):
    """BASELINE configs[3]: 512 x 512 x 64 @ 0.1 m -- two 32-cell chunks per column, 16x16-voxel coarse blocks,
    16.8 M voxels.  Small scans and two invalid passes keep the oracle at a few seconds."""
    seq = make_sequence(str(tmp_path / "seq"), 3, 32, 700)
    cnt, poses, scans, labels, raw = _load_sequence(seq)
    be = _backend("stress_512", scans, raw, frames_in_flight=1, vis_mode=vis_mode)
    assert be.size == (512, 512, 64)
    o = oracle_frame(orc, cfg_path("stress_512"), scans, labels, poses, [0], [1, 2])
    p = product_frame(be, poses, [0], [1, 2])
    assert_frames_equal(p, o)
    be.close()


def test_more_passes_than_memo_epochs(orc, tmp_path, vis_mode):
    """BASELINE configs[4] in small: far more updateInvalid passes than the 127 pass tags of the memo, so the tag
    counter restarts (twice) inside one frame."""
    seq = make_sequence(str(tmp_path / "seq"), 3, 16, 300)
    cnt, poses, scans, labels, raw = _load_sequence(seq)
    rng = np.random.default_rng(3)
    eps = np.column_stack([rng.uniform(-2, 28, 300), rng.uniform(-14, 14, 300), rng.uniform(-2.5, 5, 300)]).astype(np.float32)
    be = _backend("test_small", scans, raw, vis_mode=vis_mode)
    o = oracle_frame(orc, cfg_path("test_small"), scans, labels, poses, [0], [1, 2], extra_endpoints=eps)
    p = product_frame(be, poses, [0], [1, 2], extra_endpoints=eps)
    assert_frames_equal(p, o)
    be.close()


def test_fill_groups_and_batches_do_not_change_a_byte(orc, tmp_path):
    """The fill is organised by scan inside groups of `table_sets` frames and a call is cut into batches of
    `frames_in_flight` frames: whatever the grouping, every frame must come out the same, and equal to the oracle."""
    seq = make_sequence(str(tmp_path / "seq"), 14, 32, 500)
    cnt, poses, scans, labels, raw = _load_sequence(seq)
    cfg = vx.Config(cfg_path("example"))
    specs, windows = [], []
    for target in range(cnt):
        pb, pe, qb, qe = vx.window(cnt, cfg.pod.priorScans, cfg.pod.pastScans, target)
        prior, past = list(range(pb, pe)), list(range(qb, qe))
        anchor = poses[prior[-1]]
        ap_p, _ = vx.frame_transforms(anchor, np.stack([poses[i] for i in prior]))
        ap_q, ep = vx.frame_transforms(anchor, np.stack([poses[i] for i in past]))
        specs.append(vx.FrameSpec(np.asarray(prior, np.uint32), ap_p, np.asarray(past, np.uint32), ap_q, ep))
        windows.append((prior, past))
    results = []
    for kw in (dict(), dict(table_sets=3), dict(table_sets=1, frames_in_flight=5), dict(table_sets=4, frames_in_flight=4)):
        be = vx.Backend.from_config(cfg, **kw)
        for t, (s, l) in enumerate(zip(scans, raw)):
            be.upload_scan(t, s, l)
        results.append(be.process(specs))
        be.close()
    for other in results[1:]:
        for a, b in zip(results[0], other):
            for k in ("label", "bin", "occluded", "invalid"):
                assert np.array_equal(a[k], b[k]), k
    for f in (0, 6, cnt - 1):
        o = oracle_frame(orc, cfg_path("example"), scans, labels, poses, *windows[f])
        a = results[1][f]
        assert np.array_equal(a["label"], o["label"]) and np.array_equal(a["bin"], o["bin"])
        assert np.array_equal(a["occluded"], o["occluded_packed"]) and np.array_equal(a["invalid"], o["invalid_packed"])


# ------------------------------------------------------------------------------------------------------
# The benchmarked shape: full 71-scan windows of 64 x 2000-point scans on the 256 x 256 x 32 grid
# ------------------------------------------------------------------------------------------------------
def _synth_scans(n, seed=1234):
    params = vx.synth_params(seed=seed)
    cap = params.n_beams * params.n_azimuth
    xyzr = np.empty((n, cap, 4), np.float32)
    lab = np.empty((n, cap), np.uint32)
    counts = vx.synth_generate_many(params, 0, n, xyzr.ctypes.data, lab.ctypes.data, cap)
    poses = vx.synth_velo_poses(params, n).astype(np.float32)
    return [xyzr[t, : counts[t]] for t in range(n)], [lab[t, : counts[t]] for t in range(n)], poses


def test_full_window_frame_of_the_benchmark(orc, vis_mode):
    """One whole frame of BASELINE configs[1] exactly as bench.py runs it: semantic_kitti.cfg, 1 + 70 scans of
    ~124 k points (seed 1234, the bench's sequence), 70 updateInvalid passes -- every observable against the oracle."""
    scans, raw, poses = _synth_scans(71)
    labels = [l & 0xFFFF for l in raw]
    be = _backend("semantic_kitti", scans, raw, frames_in_flight=2, vis_mode=vis_mode)
    assert be.size == (256, 256, 32)
    prior, past = [0], list(range(1, 71))
    o = oracle_frame(orc, cfg_path("semantic_kitti"), scans, labels, poses, prior, past)
    p = product_frame(be, poses, prior, past)
    assert_frames_equal(p, o, what=("label", "bin", "occupancy", "occluded", "invalid", "occluded_packed", "invalid_packed"))
    assert o["invalid"].sum() > 1000 and o["occluded"].sum() > 100000
    be.close()


def test_last_frame_of_a_sequence_at_full_size(orc, vis_mode):
    """KittiReader.cpp:99-100: for the last scan the past window is the scan itself -- it is inserted twice and
    updateInvalid runs once with its endpoint at the origin.  Full-size scans and grid."""
    scans, raw, poses = _synth_scans(6)
    labels = [l & 0xFFFF for l in raw]
    cfg = vx.Config(cfg_path("semantic_kitti"))
    n = len(scans)
    pb, pe, qb, qe = vx.window(n, cfg.pod.priorScans, cfg.pod.pastScans, n - 1)
    assert (pb, pe, qb, qe) == (n - 1, n, n - 1, n)
    be = _backend("semantic_kitti", scans, raw, frames_in_flight=2, vis_mode=vis_mode)
    o = oracle_frame(orc, cfg_path("semantic_kitti"), scans, labels, poses, [n - 1], [n - 1])
    p = product_frame(be, poses, [n - 1], [n - 1])
    assert np.array_equal(p["endpoints"], np.zeros((1, 3), np.float32)) or np.abs(p["endpoints"]).max() < 1e-6
    assert_frames_equal(p, o)
    be.close()


def test_dense_config_thirty_scans(tmp_path):
    """BASELINE configs[2] in small: semantic_kitti_dense.cfg (stride 1) over a 32-scan full-size sequence; the frames
    with the longest windows (31 passes) and the last frames (shrinking windows, last-scan quirk) against the oracle."""
    seq = make_sequence(str(tmp_path / "seq"), 32, 64, 2000)
    cfg = cfg_path("semantic_kitti_dense")
    st = vx.gen_data(cfg, seq, str(tmp_path / "gpu"))
    assert st["frames"] == 32
    port = Oracle("port")
    port.gen_data(cfg, seq, str(tmp_path / "head"), first_frame=0, max_frames=3)
    port.gen_data(cfg, seq, str(tmp_path / "tail"), first_frame=28, max_frames=4)
    for d in ("head", "tail"):
        names = sorted(os.listdir(tmp_path / d))
        assert len(names) == (3 if d == "head" else 4) * 4
        for f in names:
            assert open(tmp_path / d / f, "rb").read() == open(tmp_path / "gpu" / f, "rb").read(), f


def test_far_field_tiles(orc):
    """Scans whose coordinates are 1e3 .. 5e8 (max range 1e9 allows them) moved into the grid by the pose: one ulp is up
    to 32 m there, so the fill kernel's whole-tile skip test must scale its slack -- and clusters at three
    different distances share every tile."""
    rng = np.random.default_rng(5)
    mn, mx, res = (0, -25.6, -2), (51.2, 25.6, 4.4), 0.2
    n = 4096
    shifts = [np.array([5.0e8, -3.0e8, 1.0e8]), np.array([3.0e7, -2.0e7, 1.0e7]), np.array([1000.0, 1000.0, 1000.0])]
    pts = np.zeros((3 * n, 4), np.float32)
    for c, shift in enumerate(shifts):
        inside = np.column_stack([rng.uniform(mn[a], mx[a], n) for a in range(3)])
        pts[c::3, :3] = (inside + shift).astype(np.float32)
    lab = rng.integers(1, 9, 3 * n).astype(np.uint32)
    poses = []
    for shift in shifts:
        m = np.eye(4, dtype=np.float32)
        m[:3, 3] = (-shift).astype(np.float32)
        poses.append(m.T.reshape(16).copy())
    be = vx.Backend(res, mn, mx, 0.0, 1e9, hidecar=False, debug=True)
    be.upload_scan(0, pts, lab)
    g = orc.grid(res, mn, mx)
    specs = [vx.FrameSpec(np.array([0], np.uint32), np.stack([pose]), np.zeros(0, np.uint32), np.zeros((0, 16), np.float32),
                          np.zeros((0, 3), np.float32)) for pose in poses]
    be.process(specs)
    occupied = []
    for f, pose in enumerate(poses):
        g.clear()
        m = pose.reshape(4, 4).T  # column-major 16 floats -> matrix
        x, y, z = pts[:, 0], pts[:, 1], pts[:, 2]
        # p = ((m0*x + m1*y) + m2*z) + m3 per row, every operation rounded to fp32 (voxelize_utils.cpp:193)
        rows = [(((m[r, 0] * x + m[r, 1] * y).astype(np.float32) + (m[r, 2] * z).astype(np.float32)).astype(np.float32) + m[r, 3]).astype(np.float32)
                for r in range(3)]
        p4 = np.column_stack(rows + [np.ones(len(pts), np.float32)]).astype(np.float32)
        g.insert(p4, lab & 0xFFFF)
        occ = (g.to_output_order(g.counts()) > 0).astype(np.uint8)
        occupied.append(int(occ.sum()))
        assert np.array_equal(be.dump(f, vx.DUMP_OCCUPANCY), occ), f
    assert occupied[0] >= 1 and occupied[1] > 500 and occupied[2] > 3000, occupied
    be.close()


def test_async_submit_overlaps_uploads(orc, tmp_path):
    """vx_submit_frames / vx_wait: a job runs while the scans of the next job are uploaded under other ids; both jobs
    give the bytes of the synchronous call."""
    seq = make_sequence(str(tmp_path / "seq"), 10, 32, 600)
    cnt, poses, scans, labels, raw = _load_sequence(seq)
    cfg = vx.Config(cfg_path("example"))

    def specs_for(id_offset):
        out = []
        for target in range(cnt):
            pb, pe, qb, qe = vx.window(cnt, cfg.pod.priorScans, cfg.pod.pastScans, target)
            anchor = poses[pe - 1]
            ap_p, _ = vx.frame_transforms(anchor, poses[pb:pe])
            ap_q, ep = vx.frame_transforms(anchor, poses[qb:qe])
            out.append(vx.FrameSpec(np.arange(pb, pe, dtype=np.uint32) + id_offset, ap_p, np.arange(qb, qe, dtype=np.uint32) + id_offset, ap_q, ep))
        return out

    be = vx.Backend.from_config(cfg, frames_in_flight=4)
    for t in range(cnt):
        be.upload_scan(t, scans[t], raw[t])
    sync_out = be.process(specs_for(0))
    out_a, out_b = be.new_outputs(cnt), be.new_outputs(cnt)
    arr_a, keep_a = be.build_descs(specs_for(0), out_a)
    ta = be.submit_descs(arr_a, cnt)
    del arr_a, keep_a                                   # the library copied the descriptors
    for t in range(cnt):                                # meanwhile: the same scans again under other ids
        be.upload_scan(1000 + t, scans[t], raw[t])
    arr_b, keep_b = be.build_descs(specs_for(1000), out_b)
    tb = be.submit_descs(arr_b, cnt)
    be.wait(ta)
    for t in range(cnt):                                # job a is done: its scans may go
        be.evict_scan(t)
    be.wait(tb)
    be.wait(0)
    for a, b, c in zip(sync_out, out_a, out_b):
        for k in ("label", "bin", "occluded", "invalid"):
            assert np.array_equal(a[k], b[k]) and np.array_equal(a[k], c[k]), k
    bad = vx.FrameSpec(np.array([77777], np.uint32), np.eye(4, dtype=np.float32).reshape(1, 16), np.zeros(0, np.uint32),
                       np.zeros((0, 16), np.float32), np.zeros((0, 3), np.float32))
    arr, keep = be.build_descs([bad], be.new_outputs(1))
    t_bad = be.submit_descs(arr, 1)
    with pytest.raises(vx.VxError, match="not resident"):
        be.wait(t_bad)
    be.close()


def test_outputs_in_a_ring_leave_as_2d_copies(tmp_path):
    """Host buffers that are evenly spaced (an output ring) are copied out with one 2-D copy per run of frames instead of four
    copies per frame: whatever the placement (in order, partly reversed, with a gap), the bytes are those of separate arrays."""
    seq = make_sequence(str(tmp_path / "seq"), 12, 32, 500)
    cnt, poses, scans, labels, raw = _load_sequence(seq)
    cfg = vx.Config(cfg_path("example"))
    specs = []
    for target in range(cnt):
        pb, pe, qb, qe = vx.window(cnt, cfg.pod.priorScans, cfg.pod.pastScans, target)
        anchor = poses[pe - 1]
        ap_p, _ = vx.frame_transforms(anchor, poses[pb:pe])
        ap_q, ep = vx.frame_transforms(anchor, poses[qb:qe])
        specs.append(vx.FrameSpec(np.arange(pb, pe, dtype=np.uint32), ap_p, np.arange(qb, qe, dtype=np.uint32), ap_q, ep))
    be = vx.Backend.from_config(cfg, table_sets=5)
    for t in range(cnt):
        be.upload_scan(t, scans[t], raw[t])
    want = be.process(specs)
    per = be.label_bytes + 3 * be.packed_bytes
    for place in (list(range(cnt)), [0, 1, 2, 5, 4, 3, 6, 7, 9, 11, 13, 14]):
        ring = vx.PinnedBuffer((max(place) + 1) * per)
        ring.array[:] = 0xAB
        outs = []
        for k in place:
            b = ring.array[k * per:(k + 1) * per]
            outs.append(dict(label=b[:be.label_bytes], bin=b[be.label_bytes:be.label_bytes + be.packed_bytes],
                             occluded=b[be.label_bytes + be.packed_bytes:be.label_bytes + 2 * be.packed_bytes],
                             invalid=b[be.label_bytes + 2 * be.packed_bytes:]))
        arr, keep = be.build_descs(specs, outs)
        be.process_descs(arr, cnt)
        for f in range(cnt):
            for k in ("label", "bin", "occluded", "invalid"):
                assert np.array_equal(outs[f][k].view(np.uint8), want[f][k].view(np.uint8)), (place[f], k)
        used = set(place)
        for k in range(max(place) + 1):
            if k not in used:
                assert (ring.array[k * per:(k + 1) * per] == 0xAB).all(), "a slot no frame owns was written"
        ring.free()
    be.close()


def test_several_jobs_and_batches_in_flight(tmp_path):
    """The library keeps `batches_in_flight` batches (lanes) in flight: five jobs of three batches each, submitted back to back
    under two scan id ranges, give the bytes of the synchronous call -- also when the histogram tables are far too small, so that
    an overflow is met while other batches are in flight (everything in flight is redone with grown tables, oldest first)."""
    seq = make_sequence(str(tmp_path / "seq"), 9, 32, 600)
    cnt, poses, scans, labels, raw = _load_sequence(seq)
    cfg = vx.Config(cfg_path("example"))

    def specs_for(id_offset, order):
        out = []
        for target in order:
            pb, pe, qb, qe = vx.window(cnt, cfg.pod.priorScans, cfg.pod.pastScans, target)
            anchor = poses[pe - 1]
            ap_p, _ = vx.frame_transforms(anchor, poses[pb:pe])
            ap_q, ep = vx.frame_transforms(anchor, poses[qb:qe])
            out.append(vx.FrameSpec(np.arange(pb, pe, dtype=np.uint32) + id_offset, ap_p, np.arange(qb, qe, dtype=np.uint32) + id_offset, ap_q, ep))
        return out

    ref = vx.Backend.from_config(cfg)
    for t in range(cnt):
        ref.upload_scan(t, scans[t], raw[t])
    want = ref.process(specs_for(0, range(cnt)))
    ref.close()
    for kw in (dict(frames_in_flight=3), dict(frames_in_flight=3, table_log2_capacity=10)):
        be = vx.Backend.from_config(cfg, **kw)
        assert be.batches_in_flight >= 2
        for t in range(cnt):
            be.upload_scan(t, scans[t], raw[t])
            be.upload_scan(500 + t, scans[t], raw[t])
        jobs = []
        for j in range(5):
            order = list(range(cnt)) if j % 2 == 0 else list(range(cnt - 1, -1, -1))
            outs = be.new_outputs(cnt)
            arr, keep = be.build_descs(specs_for(500 * (j % 2), order), outs)
            jobs.append((be.submit_descs(arr, cnt), order, outs))
            del arr, keep
        be.wait(jobs[2][0])
        be.wait(0)
        for _, order, outs in jobs:
            for target, o in zip(order, outs):
                for k in ("label", "bin", "occluded", "invalid"):
                    assert np.array_equal(o[k], want[target][k]), (kw, target, k)
        be.close()


def test_thousand_frame_job_is_deterministic():
    """The benchmarked launch shape (>= 1036 frames = 7 CTAs on each of 148 SMs, RED.MAX / RED.OR marks racing in L2):
    the same job twice gives one sha256 over every output byte; set relations hold on a sample of frames."""
    import hashlib
    n_scans = 1040 + 14
    params = vx.synth_params(seed=77, n_beams=32, n_azimuth=1000)     # 32 k-point scans keep the host side short
    cap = params.n_beams * params.n_azimuth
    xyzr = np.empty((n_scans, cap, 4), np.float32)
    lab = np.empty((n_scans, cap), np.uint32)
    counts = vx.synth_generate_many(params, 0, n_scans, xyzr.ctypes.data, lab.ctypes.data, cap)
    poses = vx.synth_velo_poses(params, n_scans).astype(np.float32)
    cfg = vx.Config(cfg_path("semantic_kitti_dense"))
    be = vx.Backend(cfg.pod.voxelSize, list(cfg.pod.minExtent)[:3], list(cfg.pod.maxExtent)[:3], cfg.pod.minRange, cfg.pod.maxRange,
                    vis_mode=1)  # (left to itself the library would give a batch just above 1036 frames the middle shape)
    auto = vx.Backend(cfg.pod.voxelSize, list(cfg.pod.minExtent)[:3], list(cfg.pod.maxExtent)[:3], cfg.pod.minRange, cfg.pod.maxRange)
    for t in range(n_scans):
        be.upload_scan(t, xyzr[t, : counts[t]], lab[t, : counts[t]])
        auto.upload_scan(t, xyzr[t, : counts[t]], lab[t, : counts[t]])
    specs = []
    for target in range(1040):
        past = np.arange(target + 1, target + 15, dtype=np.uint32)    # 14 past scans: 14 invalid passes per frame
        anchor = poses[target]
        ap_p, _ = vx.frame_transforms(anchor, poses[target:target + 1])
        ap_q, ep = vx.frame_transforms(anchor, poses[past])
        specs.append(vx.FrameSpec(np.array([target], np.uint32), ap_p, past, ap_q, ep))
    assert be.frames_in_flight >= 1040
    outs = be.new_outputs(len(specs))
    digests = []
    for _ in range(2):
        be.process(specs, outs)
        h = hashlib.sha256()
        for o in outs:
            for k in ("bin", "label", "occluded", "invalid"):
                h.update(o[k].tobytes())
        digests.append(h.hexdigest())
    assert digests[0] == digests[1]
    # the first 560 frames alone: a batch of 2 - 4 frames per SM runs in the MIDDLE shape (4 x 256-thread CTAs per SM, all of them
    # resident at once) -- same bytes as the throughput shape gave above
    part = auto.process(specs[:560])
    for a, b in zip(part, outs[:560]):
        for k in ("bin", "label", "occluded", "invalid"):
            assert np.array_equal(a[k], b[k]), k
    del part
    auto.close()
    for o in outs[::97]:
        occl = np.unpackbits(o["occluded"], bitorder="big").astype(bool)
        inv = np.unpackbits(o["invalid"], bitorder="big").astype(bool)
        assert not (inv & (o["label"] > 0)).any() and not (inv & ~occl).any() and ((o["label"] > 0) <= occl).all()
        assert inv.sum() > 0
    be.close()


def test_frame_flag_insert_occlusion_labels(orc, tmp_path, vis_mode):
    """VX_FRAME_INSERT_OCCLUSION_LABELS through the C ABI on a real two-grid frame: the target grid is filled between
    updateOcclusions and the updateInvalid passes (gen_data.cpp:108-109), the input grid's `.bin` is left as inserted."""
    seq = make_sequence(str(tmp_path / "seq"), 6, 32, 600)
    cnt, poses, scans, labels, raw = _load_sequence(seq)
    be = vx.Backend.from_config(vx.Config(cfg_path("example")), debug=True, vis_mode=vis_mode, occlusion_labels=True)
    for t, (s_, l_) in enumerate(zip(scans, raw)):
        be.upload_scan(t, s_, l_)
    o = oracle_frame(orc, cfg_path("example"), scans, labels, poses, [1], [2, 3, 4, 5], insert_occlusion_labels=True)
    p = product_frame(be, poses, [1], [2, 3, 4, 5], flags=vx.FRAME_INSERT_OCCLUSION_LABELS)
    assert_frames_equal(p, o, what=("label", "bin", "occluded", "invalid", "occluded_packed", "invalid_packed"))
    plain = oracle_frame(orc, cfg_path("example"), scans, labels, poses, [1], [2, 3, 4, 5])
    assert (o["label"] != plain["label"]).sum() > 100            # the flag did something
    src = be.dump(0, vx.DUMP_FILL_SOURCE)
    assert int((src != 0xFFFFFFFF).sum()) == int(o["occupancy"].sum()) - int(plain["occupancy"].sum())
    # without the option the flag is refused
    be2 = vx.Backend.from_config(vx.Config(cfg_path("example")))
    for t, (s_, l_) in enumerate(zip(scans, raw)):
        be2.upload_scan(t, s_, l_)
    with pytest.raises(vx.VxError, match="occlusion_labels"):
        product_frame(be2, poses, [1], [2, 3, 4, 5], flags=vx.FRAME_INSERT_OCCLUSION_LABELS)
    be.close()
    be2.close()
