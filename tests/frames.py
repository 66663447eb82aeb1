"""Frame-level helpers shared by the GPU parity tests: the same frame run through the product's C ABI and
through the oracle, with every observable compared in the OUTPUT index order L = (x*Sy + y)*Sz + z."""
from __future__ import annotations

import numpy as np

import voxelizer_b200 as vx
from oracle.oracle import pack_msb


def oracle_frame(orc, cfg_file, scans, labels, poses, prior_idx, past_idx, extra_endpoints=None, insert_occlusion_labels=False):
    """The loop body of gen_data.cpp:89-121 with the oracle; labels must already be & 0xFFFF."""
    cfg = orc.load_config(cfg_file)
    p = cfg.pod
    anchor = poses[prior_idx[-1]]
    gi = orc.grid(p.voxelSize, list(p.minExtent), list(p.maxExtent))
    gt = orc.grid(p.voxelSize, list(p.minExtent), list(p.maxExtent))
    pick = lambda idx: ([scans[i] for i in idx], [labels[i] for i in idx], np.stack([poses[i] for i in idx]) if len(idx) else np.zeros((0, 16), np.float32))
    ps, pl, pp = pick(prior_idx)
    qs, ql, qp = pick(past_idx)
    gi.fill(cfg, anchor, ps, pl, pp)
    gt.fill(cfg, anchor, ps, pl, pp)
    if len(past_idx):
        gt.fill(cfg, anchor, qs, ql, qp)
    gi.update_occlusions()
    gt.update_occlusions()
    if insert_occlusion_labels:  # gen_data.cpp:108-109 (commented out there), target grid only: VX_FRAME_INSERT_OCCLUSION_LABELS
        gt.insert_occlusion_labels()
    eps = [orc.endpoint(anchor, poses[i]) for i in past_idx]
    if extra_endpoints is not None:
        eps = list(extra_endpoints)
    for e in eps:
        gt.update_invalid(e)
    occl, inv = gt.flags()

    def majority(g):
        h = g.hist()  # (internal voxel, label, count) sorted by voxel, label
        lab = np.zeros(g.nvox, np.uint32)
        best = np.zeros(g.nvox, np.uint32)
        for v, l, c in h:  # ascending label within a voxel: strict '>' keeps the smallest label on ties
            if c > best[v]:
                best[v] = c
                lab[v] = l
        return g.to_output_order(lab).astype(np.uint16)

    lt, li = majority(gt), majority(gi)

    def hist_out(g):
        h = g.hist().astype(np.int64)
        sx, sy, sz = g.size
        v = h[:, 0]
        i, j, k = v % sx, (v // sx) % sy, v // (sx * sy)
        L = (i * sy + j) * sz + k
        out = np.column_stack([L, h[:, 1], h[:, 2]])
        return out[np.lexsort((out[:, 1], out[:, 0]))]

    return dict(
        size=gt.size, offset=gt.offset, label=lt, bin=pack_msb(li > 0), occupancy=(gt.to_output_order(gt.counts()) > 0).astype(np.uint8),
        occluded=occl, invalid=inv, occluded_packed=pack_msb(occl), invalid_packed=pack_msb(inv), hist_target=hist_out(gt),
        hist_input=hist_out(gi), endpoints=np.array(eps, np.float32).reshape(-1, 3),
    )


def product_frame(backend: vx.Backend, poses, prior_idx, past_idx, extra_endpoints=None, frame_index=0, also=None, flags=0):
    """Same frame through vx_process_frames; scans must be resident under their index as id."""
    anchor = poses[prior_idx[-1]]
    ap_p, _ = vx.frame_transforms(anchor, np.stack([poses[i] for i in prior_idx]))
    if len(past_idx):
        ap_q, ep = vx.frame_transforms(anchor, np.stack([poses[i] for i in past_idx]))
    else:
        ap_q, ep = np.zeros((0, 16), np.float32), np.zeros((0, 3), np.float32)
    if extra_endpoints is not None:
        ep = np.asarray(extra_endpoints, np.float32).reshape(-1, 3)
    spec = vx.FrameSpec(np.asarray(prior_idx, np.uint32), ap_p, np.asarray(past_idx, np.uint32), ap_q, ep, flags)
    frames = [spec] + list(also or [])
    outs = backend.process(frames)
    o = outs[frame_index]

    def sort_hist(h):
        h = h.astype(np.int64)
        return h[np.lexsort((h[:, 1], h[:, 0]))] if len(h) else h.reshape(0, 3)

    res = dict(label=o["label"], bin=o["bin"], occluded_packed=o["occluded"], invalid_packed=o["invalid"], endpoints=ep)
    res["occupancy"] = backend.dump(frame_index, vx.DUMP_OCCUPANCY)
    res["occluded"] = backend.dump(frame_index, vx.DUMP_OCCLUDED)
    res["invalid"] = backend.dump(frame_index, vx.DUMP_INVALID)
    try:
        res["hist_target"] = sort_hist(backend.dump(frame_index, vx.DUMP_HIST_TARGET))
        res["hist_input"] = sort_hist(backend.dump(frame_index, vx.DUMP_HIST_INPUT))
    except vx.VxError:
        pass
    res["all_outputs"] = outs
    return res


def assert_frames_equal(prod: dict, orc: dict, what=("hist_target", "hist_input", "label", "bin", "occupancy", "occluded", "invalid",
                                                    "occluded_packed", "invalid_packed")):
    for k in what:
        if k not in prod:
            continue
        a, b = np.asarray(prod[k]), np.asarray(orc[k])
        assert a.shape == b.shape, (k, a.shape, b.shape)
        if not np.array_equal(a, b):
            bad = np.flatnonzero((a != b).reshape(len(a), -1).any(axis=1))
            raise AssertionError(f"{k}: {len(bad)} mismatching entries, first at {bad[:5]}: {a[bad[:3]]} vs {b[bad[:3]]}")
