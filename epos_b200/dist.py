"""Multi-GPU plumbing (new relative to the reference, which is single-GPU batch-1: scripts/infer.py:610):
images are sharded across ranks (one process per GPU), weights are replicated by ONE broadcast of a packed
f32 blob at start-up, and pose records are exchanged by ONE all-gather per batch.  No other collective is on
the data path.  Backend: NCCL over NVLink on GPUs, gloo in the CPU tests."""
import numpy as np
import torch
import torch.distributed as dist

from . import weights as W


def shard_range(n_items, world, rank):
    """Contiguous shard [lo, hi) of rank `rank`: first (n % world) ranks get one extra item."""
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def pack_weights(weights, num_objs, num_frags, model_variant='xception_65'):
    """dict -> one contiguous f32 vector in variable_specs order (BN scopes expand to their 4 vectors)."""
    parts = []
    for name, shape, init, _ in W.variable_specs(num_objs, num_frags, model_variant):
        if init == 'bn':
            for k in W.BN_KEYS:
                parts.append(np.asarray(weights['%s/%s' % (name, k)], np.float32).ravel())
        else:
            a = np.asarray(weights[name], np.float32)
            assert a.shape == tuple(shape), (name, a.shape, shape)
            parts.append(a.ravel())
    return np.concatenate(parts)


def unpack_weights(blob, num_objs, num_frags, model_variant='xception_65'):
    out, off = {}, 0
    for name, shape, init, _ in W.variable_specs(num_objs, num_frags, model_variant):
        if init == 'bn':
            for k in W.BN_KEYS:
                out['%s/%s' % (name, k)] = blob[off:off + shape[0]]
                off += shape[0]
        else:
            n = int(np.prod(shape))
            out[name] = blob[off:off + n].reshape(shape)
            off += n
    assert off == blob.size, (off, blob.size)
    return out


def blob_size(num_objs, num_frags, model_variant='xception_65'):
    n = 0
    for _, shape, init, _ in W.variable_specs(num_objs, num_frags, model_variant):
        n += 4 * shape[0] if init == 'bn' else int(np.prod(shape))
    return n


def broadcast_weights(weights, num_objs, num_frags, device, world, rank, src=0, model_variant='xception_65'):
    """Rank `src` holds `weights`; every rank returns the same dict.  One collective."""
    if world == 1:
        return weights
    n = blob_size(num_objs, num_frags, model_variant)
    if rank == src:
        t = torch.from_numpy(pack_weights(weights, num_objs, num_frags, model_variant)).to(device)
    else:
        t = torch.empty(n, dtype=torch.float32, device=device)
    dist.broadcast(t, src=src)
    return unpack_weights(t.cpu().numpy(), num_objs, num_frags, model_variant)


def all_gather_poses(poses, world, group=None):
    """poses [b, O, 16] f64 on this rank -> [world*b, O, 16] (rank-major, i.e. global image order)."""
    if world == 1:
        return poses
    out = torch.empty((world * poses.shape[0],) + tuple(poses.shape[1:]), dtype=poses.dtype, device=poses.device)
    dist.all_gather_into_tensor(out, poses.contiguous(), group=group)
    return out
