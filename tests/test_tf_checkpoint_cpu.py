"""TensorFlow-free checkpoint reader (epos_b200/tf_checkpoint.py) against bundles written here by an independent minimal
writer of the documented format (LevelDB-style table with prefix-compressed keys, restart points, several data blocks;
BundleHeaderProto / BundleEntryProto values), including a whole random-init EPOS model."""
import os
import struct

import numpy as np
import pytest

from epos_b200 import tf_checkpoint as T, weights as W

_ENUM = {np.dtype(np.float32): 1, np.dtype(np.float64): 2, np.dtype(np.int32): 3, np.dtype(np.int64): 9}


def vi(n):
    out = bytearray()
    while True:
        b = n & 0x7F
        n >>= 7
        out.append(b | (0x80 if n else 0))
        if not n:
            return bytes(out)


def pb_varint(field, v):
    return vi(field << 3) + vi(v)


def pb_bytes(field, b):
    return vi((field << 3) | 2) + vi(len(b)) + b


def entry_proto(arr, shard, offset):
    shape = b''.join(pb_bytes(2, pb_varint(1, d)) for d in arr.shape)
    return (pb_varint(1, _ENUM[arr.dtype]) + pb_bytes(2, shape) + (pb_varint(3, shard) if shard else b'') +
            pb_varint(4, offset) + pb_varint(5, arr.nbytes) + vi((6 << 3) | 5) + struct.pack('<I', 0xdeadbeef))


def build_block(items, restart_interval=3):
    buf, restarts, prev = bytearray(), [], b''
    for i, (k, v) in enumerate(items):
        shared = 0
        if i % restart_interval == 0:
            restarts.append(len(buf))
        else:
            while shared < min(len(prev), len(k)) and prev[shared] == k[shared]:
                shared += 1
        buf += vi(shared) + vi(len(k) - shared) + vi(len(v)) + k[shared:] + v
        prev = k
    for r in restarts or [0]:
        buf += struct.pack('<I', r)
    buf += struct.pack('<I', max(1, len(restarts)))
    return bytes(buf)


def write_bundle(prefix, tensors, num_shards=1, per_block=5):
    """tensors: {name: array}.  Tensors are spread round-robin over the shards."""
    names = sorted(tensors)
    data = [bytearray() for _ in range(num_shards)]
    items = [(b'', pb_varint(1, num_shards) + pb_varint(2, 0))]
    for i, n in enumerate(names):
        a = np.asarray(tensors[n], order='C')          # (ascontiguousarray would promote scalars to 1-d)
        s = i % num_shards
        items.append((n.encode(), entry_proto(a, s, len(data[s]))))
        data[s] += a.tobytes()
    out, handles = bytearray(), []
    for b0 in range(0, len(items), per_block):
        blk = build_block(items[b0:b0 + per_block])
        handles.append((items[min(b0 + per_block, len(items)) - 1][0], len(out), len(blk)))
        out += blk + b'\x00' + struct.pack('<I', 0)          # trailer: no compression, crc (unchecked)
    meta = build_block([])
    meta_h = (len(out), len(meta)); out += meta + b'\x00' + struct.pack('<I', 0)
    idx = build_block([(k, vi(o) + vi(s)) for k, o, s in handles], restart_interval=1)
    idx_h = (len(out), len(idx)); out += idx + b'\x00' + struct.pack('<I', 0)
    footer = vi(meta_h[0]) + vi(meta_h[1]) + vi(idx_h[0]) + vi(idx_h[1])
    out += footer + b'\x00' * (40 - len(footer)) + struct.pack('<Q', T.TABLE_MAGIC)
    with open(prefix + '.index', 'wb') as f:
        f.write(out)
    for s in range(num_shards):
        with open('%s.data-%05d-of-%05d' % (prefix, s, num_shards), 'wb') as f:
            f.write(data[s])


def test_round_trip_dtypes_shapes_blocks_and_shards(tmp_path):
    rng = np.random.default_rng(0)
    t = {'a/weights': rng.standard_normal((3, 3, 4, 5)).astype(np.float32),
         'a/BatchNorm/gamma': rng.standard_normal(5).astype(np.float32),
         'a/BatchNorm/beta': rng.standard_normal(5).astype(np.float32),
         'global_step': np.array(123456, np.int64),
         'b/scalar64': np.array(2.5, np.float64),
         'b/ints': np.arange(12, dtype=np.int32).reshape(3, 4),
         'zzz/long/name/with/a/common/prefix/one': rng.standard_normal((7,)).astype(np.float32),
         'zzz/long/name/with/a/common/prefix/two': rng.standard_normal((2, 9)).astype(np.float32)}
    for shards in (1, 3):
        prefix = str(tmp_path / ('m%d.ckpt-7' % shards))
        write_bundle(prefix, t, num_shards=shards, per_block=3)
        got = T.load_checkpoint(prefix)
        assert sorted(got) == sorted(t)
        for k in t:
            assert got[k].dtype == t[k].dtype and got[k].shape == t[k].shape and np.array_equal(got[k], t[k])
        assert [n for n, _, _ in T.list_variables(prefix)] == sorted(t)
        sub = T.load_checkpoint(prefix, ['b/ints'])
        assert list(sub) == ['b/ints']
        with pytest.raises(KeyError):
            T.load_checkpoint(prefix, ['missing'])


def test_rejects_non_checkpoints(tmp_path):
    p = str(tmp_path / 'x')
    with open(p + '.index', 'wb') as f:
        f.write(b'\x00' * 100)
    with pytest.raises(ValueError):
        T.read_index(p + '.index')


@pytest.mark.parametrize('variant', ['xception_65', 'resnet_v1_50_beta'])
def test_whole_epos_model_with_optimizer_slots(tmp_path, variant):
    O, F = 2, 4
    w = W.random_init(O, F, seed=3, bn='random', model_variant=variant)
    extra = {k + '/Momentum': np.zeros_like(v) for k, v in list(w.items())[:5]}
    extra['global_step'] = np.array(10, np.int64)
    prefix = str(tmp_path / 'model.ckpt-10')
    write_bundle(prefix, dict(w, **extra), per_block=16)
    got, o, f = T.epos_weights_from_checkpoint(prefix, variant)
    assert (o, f) == (O, F) and sorted(got) == sorted(w)
    for k in w:
        assert np.array_equal(got[k], w[k])
    out = str(tmp_path / 'w.npz')
    W.save_npz(out, got)
    back = W.load_npz(out)
    assert all(np.array_equal(back[k], w[k]) for k in w)
