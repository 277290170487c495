"""world_size-2 gloo tests (CPU) of the multi-GPU plumbing: one broadcast of the packed weight blob, image sharding,
one all-gather of pose records per batch (epos_b200/dist.py)."""
import json
import os
import socket
import subprocess
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_broadcast_shard_gather_world2(tmp_path):
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2', '--master-addr', '127.0.0.1',
           '--master-port', str(_free_port()), os.path.join(HERE, 'dist_worker.py'), str(tmp_path)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=240)
    assert r.returncode == 0, r.stderr[-2000:]
    res = [json.load(open(os.path.join(tmp_path, 'rank%d.json' % k))) for k in range(2)]
    assert [tuple(x['shard']) for x in res] == [(0, 3), (3, 5)]
    for x in res:
        assert x['ok_w'] and x['shape'] == [6, 3, 16]
        assert x['vals'] == [1.0, 2.0, 3.0, 4.0, 5.0, 0.0]     # rank-major global image order, zero padding at the tail


def test_world1_is_a_passthrough():
    from epos_b200 import dist as ed
    t = torch.ones((2, 3, 16), dtype=torch.float64)
    assert ed.all_gather_poses(t, 1) is t
    w = {'a': np.zeros(3)}
    assert ed.broadcast_weights(w, 1, 1, torch.device('cpu'), 1, 0) is w
