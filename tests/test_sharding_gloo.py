"""N>1 host logic on CPU: two ranks (gloo) shard the target list as bench.py / vx_gen_data --shard do --
contiguous slices, no data-path collective, one gather of per-rank timings."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import voxelizer_b200 as vx
from util import cfg_path, make_sequence


def shard_bounds(total: int, rank: int, world: int):
    return total * rank // world, total * (rank + 1) // world


def _worker(rank, world, port, seq, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    cfg = vx.Config(cfg_path("semantic_kitti_dense"))
    cnt, poses = vx.read_poses(seq)
    targets = vx.plan_targets(cfg, cnt, poses)
    lo, hi = shard_bounds(len(targets), rank, world)
    mine = targets[lo:hi]
    # halo: the scans this rank must hold
    need = set()
    for t in mine:
        pb, pe, qb, qe = vx.window(cnt, cfg.pod.priorScans, 4, int(t))
        need.update(range(pb, pe))
        need.update(range(qb, qe))
    t_local = torch.tensor([float(10 + rank)], dtype=torch.float64)  # stand-in for a rank's device time
    gathered = [torch.zeros(1, dtype=torch.float64) for _ in range(world)]
    dist.all_gather(gathered, t_local)
    count = torch.tensor([len(mine)], dtype=torch.int64)
    dist.all_reduce(count)
    q.put((rank, mine.tolist(), sorted(need), [float(g) for g in gathered], int(count)))
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_partition_the_targets(tmp_path):
    seq = make_sequence(str(tmp_path / "seq"), 21, 8, 64)
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, seq, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, m0, need0, g0, c0), (r1, m1, need1, g1, c1) = res
    assert m0 + m1 == list(range(21))            # contiguous, disjoint, complete
    assert abs(len(m0) - len(m1)) <= 1
    assert c0 == c1 == 21
    assert g0 == g1 == [10.0, 11.0] and max(g0) == 11.0   # every rank sees all timings; the job time is the max
    assert need0[0] == 0 and need0[-1] == m0[-1] + 4      # halo of `past` scans beyond the slice
    assert need1[0] == m1[0] and need1[-1] == 20


def test_shard_bounds_cover_everything():
    for total in (0, 1, 7, 909, 4541):
        for world in (1, 2, 4, 8):
            b = [shard_bounds(total, r, world) for r in range(world)]
            assert b[0][0] == 0 and b[-1][1] == total
            assert all(b[i][1] == b[i + 1][0] for i in range(world - 1))
