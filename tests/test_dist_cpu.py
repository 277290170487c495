"""world_size-2 gloo tests (CPU) of the multi-GPU plumbing: one broadcast of the packed weight blob, image sharding,
one all-gather of pose records per batch (epos_b200/dist.py)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        from epos_b200 import dist as ed, weights as W
        O, F = 2, 4
        w = W.random_init(O, F, seed=7, bn='random') if rank == 0 else None
        got = ed.broadcast_weights(w, O, F, torch.device('cpu'), world, rank)
        ref = W.random_init(O, F, seed=7, bn='random')
        ok_w = set(got) == set(ref) and all(np.array_equal(got[k], ref[k]) for k in ref)
        # shards of a 5-image batch and the gathered pose records
        n_img, J = 5, 3
        lo, hi = ed.shard_range(n_img, world, rank)
        # equal-size shards are required by all_gather_into_tensor: pad to the largest shard
        per = max(ed.shard_range(n_img, world, r)[1] - ed.shard_range(n_img, world, r)[0] for r in range(world))
        mine = torch.zeros((per, J, 16), dtype=torch.float64)
        for i in range(lo, hi):
            mine[i - lo] = float(i + 1)
        allp = ed.all_gather_poses(mine, world)
        ok_g = tuple(allp.shape) == (world * per, J, 16)
        vals = [float(allp[r * per + k, 0, 0]) for r in range(world) for k in range(per)]
        q.put((rank, ok_w, ok_g, (lo, hi), vals))
    finally:
        dist.destroy_process_group()


def test_broadcast_shard_gather_world2():
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert [r[3] for r in res] == [(0, 3), (3, 5)]
    for rank, ok_w, ok_g, _, vals in res:
        assert ok_w and ok_g
        assert vals == [1.0, 2.0, 3.0, 4.0, 5.0, 0.0]       # rank-major global image order, zero padding at the tail


def test_world1_is_a_passthrough():
    from epos_b200 import dist as ed
    t = torch.ones((2, 3, 16), dtype=torch.float64)
    assert ed.all_gather_poses(t, 1) is t
    w = {'a': np.zeros(3)}
    assert ed.broadcast_weights(w, 1, 1, torch.device('cpu'), 1, 0) is w
