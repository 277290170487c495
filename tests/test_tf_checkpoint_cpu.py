"""TensorFlow-free checkpoint reader (epos_b200/tf_checkpoint.py) against bundles written here by an independent minimal
writer of the documented format (LevelDB-style table with prefix-compressed keys, restart points, several data blocks;
BundleHeaderProto / BundleEntryProto values), including a whole random-init EPOS model.

Pins that do not come from this repository's own reading of the format:
  * CRC-32C: the RFC 3720 (iSCSI) appendix B.4 test vectors, which LevelDB's crc32c_test.cc also uses, and LevelDB's mask;
  * the checksums of the bundles below are produced by TensorBoard's implementation (tensorboard.compat.tensorflow_stub,
    the code that writes TFRecord/event files without TensorFlow), not by the reader's own;
  * BundleHeaderProto / BundleEntryProto / TensorShapeProto values are serialised by the official protobuf runtime from
    descriptors built out of the published schema (tensorflow/core/protobuf/tensor_bundle.proto, tensor_shape.proto),
    not by the hand-written encoder, in test_entries_encoded_by_the_protobuf_runtime."""
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


def tb_masked_crc(data):
    """Masked CRC-32C by TensorBoard's TensorFlow stub (independent of the reader under test)."""
    from tensorboard.compat.tensorflow_stub import pywrap_tensorflow as tb
    return int(tb.masked_crc32c(bytes(data)))


def entry_proto(arr, shard, offset, checksum=True):
    shape = b''.join(pb_bytes(2, pb_varint(1, d)) for d in arr.shape)
    crc = tb_masked_crc(arr.tobytes()) if checksum else 0
    return (pb_varint(1, _ENUM[arr.dtype]) + pb_bytes(2, shape) + (pb_varint(3, shard) if shard else b'') +
            pb_varint(4, offset) + pb_varint(5, arr.nbytes) + vi((6 << 3) | 5) + struct.pack('<I', crc))


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


def trailer(blk, checksum):
    """5-byte block trailer: compression type 0 + masked crc32c over (block, type byte); zero = writer did not checksum."""
    return b'\x00' + struct.pack('<I', tb_masked_crc(blk + b'\x00') if checksum else 0)


def write_bundle(prefix, tensors, num_shards=1, per_block=5, checksum=True, entry_fn=None):
    """tensors: {name: array}.  Tensors are spread round-robin over the shards."""
    entry_fn = entry_fn or (lambda a, s_, o: entry_proto(a, s_, o, checksum))
    names = sorted(tensors)
    data = [bytearray() for _ in range(num_shards)]
    items = [(b'', pb_varint(1, num_shards) + pb_varint(2, 0))]
    for i, n in enumerate(names):
        a = np.asarray(tensors[n], order='C')          # (ascontiguousarray would promote scalars to 1-d)
        s = i % num_shards
        items.append((n.encode(), entry_fn(a, s, len(data[s]))))
        data[s] += a.tobytes()
    out, handles = bytearray(), []
    for b0 in range(0, len(items), per_block):
        blk = build_block(items[b0:b0 + per_block])
        handles.append((items[min(b0 + per_block, len(items)) - 1][0], len(out), len(blk)))
        out += blk + trailer(blk, checksum)
    meta = build_block([])
    meta_h = (len(out), len(meta)); out += meta + trailer(meta, checksum)
    idx = build_block([(k, vi(o) + vi(s)) for k, o, s in handles], restart_interval=1)
    idx_h = (len(out), len(idx)); out += idx + trailer(idx, checksum)
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
    write_bundle(prefix, dict(w, **extra), per_block=16, checksum=False)      # 160 MB: TensorBoard's CRC is pure Python
    got, o, f = T.epos_weights_from_checkpoint(prefix, variant)
    assert (o, f) == (O, F) and sorted(got) == sorted(w)
    for k in w:
        assert np.array_equal(got[k], w[k])
    out = str(tmp_path / 'w.npz')
    W.save_npz(out, got)
    back = W.load_npz(out)
    assert all(np.array_equal(back[k], w[k]) for k in w)


def test_crc32c_rfc3720_vectors_and_leveldb_mask():
    """RFC 3720 appendix B.4 (the vectors of leveldb/util/crc32c_test.cc) and the LevelDB mask (crc32c.h)."""
    assert T.crc32c(b'\x00' * 32) == 0x8a9136aa
    assert T.crc32c(b'\xff' * 32) == 0x62a8ab43
    assert T.crc32c(bytes(range(32))) == 0x46dd794e
    assert T.crc32c(bytes(range(31, -1, -1))) == 0x113fdb5c
    iscsi_read = bytes([0x01, 0xc0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0x14, 0, 0, 0, 0, 0, 0x04, 0, 0, 0, 0, 0x14,
                        0, 0, 0, 0x18, 0x28, 0, 0, 0, 0, 0, 0, 0, 0x02, 0, 0, 0, 0, 0, 0, 0])
    assert T.crc32c(iscsi_read) == 0xd9963a56
    assert T.crc32c(b'123456789') == 0xe3069283                                   # the CRC catalogue's check value
    assert T.crc32c(b'world', T.crc32c(b'hello ')) == T.crc32c(b'hello world')    # Extend
    c = T.crc32c(b'foo')
    assert T.mask_crc(c) != c and T.unmask_crc(T.mask_crc(c)) == c and T.unmask_crc(T.unmask_crc(T.mask_crc(T.mask_crc(c)))) == c
    assert T.mask_crc(c) == ((((c >> 15) | (c << 17)) + 0xa282ead8) & 0xffffffff)
    rng = np.random.default_rng(0)
    for n in (0, 1, 7, 64, 1000):
        b = rng.integers(0, 256, n, dtype=np.uint8).tobytes()
        assert T.mask_crc(T.crc32c(b)) == tb_masked_crc(b)                        # TensorBoard's implementation


def test_checksums_are_verified(tmp_path):
    rng = np.random.default_rng(1)
    tensors = {'a/w': rng.normal(size=(3, 4)).astype(np.float32), 'b': np.arange(5, dtype=np.int64)}
    p = str(tmp_path / 'ok')
    write_bundle(p, tensors)
    got = T.load_checkpoint(p, verify_data=True)
    assert np.array_equal(got['a/w'], tensors['a/w']) and np.array_equal(got['b'], tensors['b'])
    # a flipped bit in a shard: the tensor's own crc32c catches it
    raw = bytearray(open(p + '.data-00000-of-00001', 'rb').read())
    raw[5] ^= 0x10
    open(p + '.data-00000-of-00001', 'wb').write(bytes(raw))
    with pytest.raises(ValueError, match='data checksum'):
        T.load_checkpoint(p)
    # a flipped bit in the index: the table block trailer catches it
    p2 = str(tmp_path / 'idx')
    write_bundle(p2, tensors)
    idx = bytearray(open(p2 + '.index', 'rb').read())
    idx[3] ^= 0x01
    open(p2 + '.index', 'wb').write(bytes(idx))
    with pytest.raises(ValueError, match='checksum'):
        T.read_index(p2 + '.index')
    # writers that leave the checksum fields zero are accepted (nothing to verify)
    p3 = str(tmp_path / 'nocrc')
    write_bundle(p3, tensors, checksum=False)
    assert np.array_equal(T.load_checkpoint(p3)['b'], tensors['b'])


def _bundle_messages():
    """BundleHeaderProto / BundleEntryProto / TensorShapeProto message classes built by the protobuf runtime from the
    published schema (field numbers and types of tensor_bundle.proto, tensor_shape.proto, types.proto)."""
    from google.protobuf import descriptor_pb2, descriptor_pool, message_factory
    fd = descriptor_pb2.FileDescriptorProto()
    fd.name = 'epos_test_tensor_bundle.proto'
    fd.package = 'epos_test'
    fd.syntax = 'proto3'
    F = descriptor_pb2.FieldDescriptorProto

    def msg(name, fields, nested=None):
        m = fd.message_type.add() if nested is None else nested.nested_type.add()
        m.name = name
        for fname, num, ftype, label, tname in fields:
            f = m.field.add()
            f.name, f.number, f.type, f.label = fname, num, ftype, label
            if tname:
                f.type_name = tname
        return m
    shape = msg('TensorShapeProto', [('dim', 2, F.TYPE_MESSAGE, F.LABEL_REPEATED, '.epos_test.TensorShapeProto.Dim'),
                                     ('unknown_rank', 3, F.TYPE_BOOL, F.LABEL_OPTIONAL, None)])
    msg('Dim', [('size', 1, F.TYPE_INT64, F.LABEL_OPTIONAL, None), ('name', 2, F.TYPE_STRING, F.LABEL_OPTIONAL, None)], nested=shape)
    msg('BundleHeaderProto', [('num_shards', 1, F.TYPE_INT32, F.LABEL_OPTIONAL, None),
                              ('endianness', 2, F.TYPE_INT32, F.LABEL_OPTIONAL, None)])
    msg('BundleEntryProto', [('dtype', 1, F.TYPE_INT32, F.LABEL_OPTIONAL, None),
                             ('shape', 2, F.TYPE_MESSAGE, F.LABEL_OPTIONAL, '.epos_test.TensorShapeProto'),
                             ('shard_id', 3, F.TYPE_INT32, F.LABEL_OPTIONAL, None),
                             ('offset', 4, F.TYPE_INT64, F.LABEL_OPTIONAL, None),
                             ('size', 5, F.TYPE_INT64, F.LABEL_OPTIONAL, None),
                             ('crc32c', 6, F.TYPE_FIXED32, F.LABEL_OPTIONAL, None)])
    pool = descriptor_pool.DescriptorPool()
    pool.Add(fd)
    get = getattr(message_factory, 'GetMessageClass', None)
    if get is None:
        fac = message_factory.MessageFactory(pool)
        get = fac.GetPrototype
    return {n: get(pool.FindMessageTypeByName('epos_test.' + n)) for n in ('BundleHeaderProto', 'BundleEntryProto')}


def test_entries_encoded_by_the_protobuf_runtime(tmp_path):
    pytest.importorskip('google.protobuf')
    M = _bundle_messages()

    def entry_pb(arr, shard, offset):
        e = M['BundleEntryProto']()
        e.dtype = _ENUM[arr.dtype]
        for d in arr.shape:
            e.shape.dim.add().size = d
        e.shard_id, e.offset, e.size, e.crc32c = shard, offset, arr.nbytes, tb_masked_crc(arr.tobytes())
        return e.SerializeToString()
    rng = np.random.default_rng(2)
    tensors = {'xception_65/entry_flow/conv1_1/weights': rng.normal(size=(3, 3, 3, 32)).astype(np.float32),
               'global_step': np.array(123456789012, np.int64), 'logits/pred_obj_conf/biases': rng.normal(size=(22,)),
               'empty': np.zeros((0, 4), np.float32), 'big_offset': rng.integers(0, 9, (300, 70)).astype(np.int32)}
    p = str(tmp_path / 'pb')
    write_bundle(p, tensors, num_shards=2, per_block=2, entry_fn=entry_pb)
    got = T.load_checkpoint(p, verify_data=True)
    assert sorted(got) == sorted(tensors)
    for k, v in tensors.items():
        assert got[k].dtype == v.dtype and got[k].shape == v.shape and np.array_equal(got[k], v)
    hdr = M['BundleHeaderProto'](num_shards=2, endianness=0).SerializeToString()
    assert hdr == pb_varint(1, 2)           # proto3 omits the zero endianness: the reader's default must be little endian
    assert T.read_index(p + '.index')[1]['endianness'] == 0
