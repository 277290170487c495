"""Worker of tests/test_dist_cpu.py: run under torch.distributed.run with the gloo backend (CPU)."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main(out_dir):
    dist.init_process_group('gloo')
    rank, world = dist.get_rank(), dist.get_world_size()
    from epos_b200 import dist as ed, weights as W
    O, F = 2, 4
    w = W.random_init(O, F, seed=7, bn='random') if rank == 0 else None
    got = ed.broadcast_weights(w, O, F, torch.device('cpu'), world, rank)
    ref = W.random_init(O, F, seed=7, bn='random')
    ok_w = set(got) == set(ref) and all(np.array_equal(got[k], ref[k]) for k in ref)
    n_img, J = 5, 3
    lo, hi = ed.shard_range(n_img, world, rank)
    per = max(ed.shard_range(n_img, world, r)[1] - ed.shard_range(n_img, world, r)[0] for r in range(world))
    mine = torch.zeros((per, J, 16), dtype=torch.float64)      # equal-size shards: pad to the largest one
    for i in range(lo, hi):
        mine[i - lo] = float(i + 1)
    allp = ed.all_gather_poses(mine, world)
    vals = [float(allp[r * per + k, 0, 0]) for r in range(world) for k in range(per)]
    with open(os.path.join(out_dir, 'rank%d.json' % rank), 'w') as f:
        json.dump({'rank': rank, 'ok_w': bool(ok_w), 'shape': list(allp.shape), 'shard': [lo, hi], 'vals': vals}, f)
    dist.destroy_process_group()


if __name__ == '__main__':
    main(sys.argv[1])
