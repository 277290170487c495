"""TensorFlow-free reader of TF checkpoints ("tensor bundle", V2 format: `<prefix>.index` + `<prefix>.data-0000i-of-0000n`).

The reference restores its weights from such a checkpoint with `tf.train.Saver` (/root/reference/epos_lib/misc.py:159-168,
scripts/infer.py:670-683).  This module reads the same files without TensorFlow so that a published EPOS model can be
converted once into the `.npz` of TF-named variables that `epos_b200.weights.load_npz` consumes
(`scripts/convert_checkpoint.py`).

Format (tensorflow/core/util/tensor_bundle + tensorflow/core/lib/io/table, which is LevelDB's sorted-table format):

  index file  = data blocks ... | metaindex block | index block | footer (48 bytes)
  footer      = metaindex BlockHandle, index BlockHandle (varint64 offset, varint64 size each), zero padding to 40 bytes,
                magic 0xdb4775248b80fb57 (little endian)
  block       = entries, restart offsets (uint32 x n), n (uint32); followed in the file by a 5-byte trailer
                (compression type, masked crc32c).  The bundle writer disables compression (type 0).
  entry       = varint32 shared key bytes, varint32 unshared key bytes, varint32 value length, key suffix, value
  key ""      -> BundleHeaderProto {1: num_shards, 2: endianness, 3: version}
  key <name>  -> BundleEntryProto  {1: dtype, 2: TensorShapeProto {2: Dim {1: size}}, 3: shard_id, 4: offset, 5: size,
                                    6: crc32c (fixed32), 7: slices}
  data files  = raw little-endian tensor bytes at [offset, offset + size)
"""
import os
import struct

import numpy as np

TABLE_MAGIC = 0xdb4775248b80fb57
_DTYPES = {1: np.float32, 2: np.float64, 3: np.int32, 4: np.uint8, 5: np.int16, 6: np.int8, 9: np.int64, 10: np.bool_,
           17: np.uint16, 19: np.float16, 22: np.uint32, 23: np.uint64}


# CRC-32C (Castagnoli, reflected polynomial 0x82F63B78) as used by LevelDB tables and the tensor bundle, and LevelDB's
# "masked" form stored in files (leveldb/util/crc32c.h: rotate right by 15, add 0xa282ead8).
_CRC_TABLE = []
for _i in range(256):
    _c = _i
    for _ in range(8):
        _c = (_c >> 1) ^ 0x82F63B78 if _c & 1 else _c >> 1
    _CRC_TABLE.append(_c)
_CRC_VERIFY_LIMIT = 1 << 20          # tensor payloads above this size are checked only with verify_data=True (pure Python)


def crc32c(data, crc=0):
    c = crc ^ 0xFFFFFFFF
    tab = _CRC_TABLE
    for b in bytes(data):
        c = tab[(c ^ b) & 0xFF] ^ (c >> 8)
    return c ^ 0xFFFFFFFF


def mask_crc(crc):
    return ((((crc >> 15) | (crc << 17)) & 0xFFFFFFFF) + 0xA282EAD8) & 0xFFFFFFFF


def unmask_crc(masked):
    rot = (masked - 0xA282EAD8) & 0xFFFFFFFF
    return ((rot >> 17) | (rot << 15)) & 0xFFFFFFFF


def _varint(buf, pos):
    result, shift = 0, 0
    while True:
        b = buf[pos]
        pos += 1
        result |= (b & 0x7F) << shift
        if not b & 0x80:
            return result, pos
        shift += 7
        if shift > 63:
            raise ValueError('malformed varint')


def _proto_fields(buf):
    """Minimal protobuf wire-format walk: yields (field number, wire type, value)."""
    pos, n = 0, len(buf)
    while pos < n:
        tag, pos = _varint(buf, pos)
        field, wt = tag >> 3, tag & 7
        if wt == 0:
            v, pos = _varint(buf, pos)
        elif wt == 1:
            v = buf[pos:pos + 8]; pos += 8
        elif wt == 2:
            ln, pos = _varint(buf, pos)
            v = buf[pos:pos + ln]; pos += ln
        elif wt == 5:
            v = buf[pos:pos + 4]; pos += 4
        else:
            raise ValueError('unsupported protobuf wire type %d' % wt)
        yield field, wt, v


def _signed64(v):
    return v - (1 << 64) if v >= (1 << 63) else v


def _parse_shape(buf):
    dims = []
    for field, wt, v in _proto_fields(buf):
        if field == 2 and wt == 2:                       # Dim
            size = 0
            for f2, w2, v2 in _proto_fields(v):
                if f2 == 1 and w2 == 0:
                    size = _signed64(v2)
            dims.append(size)
        elif field == 3 and wt == 0 and v:
            raise ValueError('tensor of unknown rank in checkpoint')
    return tuple(dims)


def _parse_entry(buf):
    e = {'dtype': 0, 'shape': (), 'shard_id': 0, 'offset': 0, 'size': 0, 'crc32c': None, 'sliced': False}
    for field, wt, v in _proto_fields(buf):
        if field == 1 and wt == 0:
            e['dtype'] = v
        elif field == 2 and wt == 2:
            e['shape'] = _parse_shape(v)
        elif field == 3 and wt == 0:
            e['shard_id'] = v
        elif field == 4 and wt == 0:
            e['offset'] = _signed64(v)
        elif field == 5 and wt == 0:
            e['size'] = _signed64(v)
        elif field == 6 and wt == 5:
            e['crc32c'] = struct.unpack('<I', v)[0]
        elif field == 7:
            e['sliced'] = True
    return e


def _block_handle(buf, pos):
    off, pos = _varint(buf, pos)
    size, pos = _varint(buf, pos)
    return off, size, pos


def _read_block(data, offset, size):
    """Returns the (key, value) pairs of the block at [offset, offset+size) (+ 5-byte trailer)."""
    if offset + size + 5 > len(data):
        raise ValueError('block handle outside the index file')
    ctype = data[offset + size]
    if ctype != 0:
        raise NotImplementedError('compressed table block (type %d); TensorFlow writes bundle indices uncompressed' % ctype)
    blk = data[offset:offset + size]
    if size < 4:
        raise ValueError('short block')
    # block trailer: masked crc32c over the block contents and the compression-type byte (table/format.cc).  Writers
    # that do not checksum leave the field zero; anything else must match.
    stored = struct.unpack('<I', data[offset + size + 1:offset + size + 5])[0]
    if stored != 0 and unmask_crc(stored) != crc32c(data[offset:offset + size + 1]):
        raise ValueError('table block checksum mismatch at offset %d (corrupt checkpoint index)' % offset)
    n_restarts = struct.unpack('<I', blk[-4:])[0]
    limit = size - 4 - 4 * n_restarts
    if limit < 0:
        raise ValueError('bad restart array')
    out, pos, key = [], 0, b''
    while pos < limit:
        shared, pos = _varint(blk, pos)
        unshared, pos = _varint(blk, pos)
        vlen, pos = _varint(blk, pos)
        if shared > len(key):
            raise ValueError('bad key prefix length')
        key = key[:shared] + bytes(blk[pos:pos + unshared]); pos += unshared
        out.append((key, bytes(blk[pos:pos + vlen]))); pos += vlen
    return out


def read_index(index_path):
    """{tensor name: entry dict} and the header dict of `<prefix>.index`."""
    with open(index_path, 'rb') as f:
        data = f.read()
    if len(data) < 48 or struct.unpack('<Q', data[-8:])[0] != TABLE_MAGIC:
        raise ValueError('%s is not a TensorFlow checkpoint index (bad table magic)' % index_path)
    footer = data[-48:]
    _, _, pos = _block_handle(footer, 0)                 # metaindex (unused)
    idx_off, idx_size, _ = _block_handle(footer, pos)
    entries, header = {}, {'num_shards': 1, 'endianness': 0}
    for _, handle in _read_block(data, idx_off, idx_size):
        off, size, _ = _block_handle(handle, 0)
        for key, value in _read_block(data, off, size):
            if key == b'':
                for field, wt, v in _proto_fields(value):
                    if field == 1 and wt == 0:
                        header['num_shards'] = v
                    elif field == 2 and wt == 0:
                        header['endianness'] = v
            else:
                entries[key.decode('utf-8')] = _parse_entry(value)
    if header['endianness'] != 0:
        raise NotImplementedError('big-endian checkpoint')
    return entries, header


def list_variables(prefix):
    """[(name, shape, numpy dtype)] like tf.train.list_variables."""
    entries, _ = read_index(prefix + '.index')
    return [(k, e['shape'], _DTYPES.get(e['dtype'])) for k, e in sorted(entries.items())]


def load_checkpoint(prefix, names=None, verify_data=False):
    """Reads the tensors of the checkpoint `prefix` (all of them, or those in `names`) into {name: numpy array}.
    The masked crc32c the bundle stores per tensor (BundleEntryProto.crc32c) is verified for payloads up to 1 MB, for
    every tensor with verify_data=True (pure-Python CRC: about 1 MB/s)."""
    entries, header = read_index(prefix + '.index')
    n = header['num_shards']
    wanted = sorted(entries) if names is None else list(names)
    files, out = {}, {}
    try:
        for name in wanted:
            if name not in entries:
                raise KeyError('variable %r is not in the checkpoint' % name)
            e = entries[name]
            if e['sliced']:
                raise NotImplementedError('partitioned variable %r' % name)
            dt = _DTYPES.get(e['dtype'])
            if dt is None:
                raise NotImplementedError('dtype enum %d of %r' % (e['dtype'], name))
            sid = e['shard_id']
            if sid not in files:
                files[sid] = open('%s.data-%05d-of-%05d' % (prefix, sid, n), 'rb')
            f = files[sid]
            f.seek(e['offset'])
            raw = f.read(e['size'])
            count = int(np.prod(e['shape'])) if e['shape'] else 1
            if len(raw) != e['size'] or e['size'] != count * np.dtype(dt).itemsize:
                raise ValueError('size mismatch for %r: %d bytes for shape %s' % (name, e['size'], (e['shape'],)))
            if e['crc32c'] not in (None, 0) and (verify_data or len(raw) <= _CRC_VERIFY_LIMIT):
                if unmask_crc(e['crc32c']) != crc32c(raw):
                    raise ValueError('tensor %r: data checksum mismatch (corrupt checkpoint shard)' % name)
            out[name] = np.frombuffer(raw, dtype=np.dtype(dt).newbyteorder('<')).reshape(e['shape']).astype(dt)
    finally:
        for f in files.values():
            f.close()
    return out


def epos_weights_from_checkpoint(prefix, model_variant='xception_65'):
    """The inference variables of an EPOS checkpoint as the dict `epos_b200.model.EposNet` takes, plus (num_objs, num_frags
    hint).  Optimizer slots, global_step and EMA shadows are ignored; missing variables raise KeyError."""
    from . import weights as W
    entries, _ = read_index(prefix + '.index')
    oc = entries.get('logits/pred_obj_conf/weights')
    fc = entries.get('logits/pred_frag_conf/weights')
    if oc is None or fc is None:
        raise KeyError('logits/pred_obj_conf/weights or logits/pred_frag_conf/weights not found: not an EPOS checkpoint')
    num_objs = oc['shape'][-1] - 1
    if num_objs <= 0 or fc['shape'][-1] % num_objs:
        raise ValueError('inconsistent head shapes %s / %s' % (oc['shape'], fc['shape']))
    num_frags = fc['shape'][-1] // num_objs
    names = []
    for name, shape, init, _ in W.variable_specs(num_objs, num_frags, model_variant):
        names += ['%s/%s' % (name, k) for k in W.BN_KEYS] if init == 'bn' else [name]
    w = load_checkpoint(prefix, names)
    for name, shape, init, _ in W.variable_specs(num_objs, num_frags, model_variant):
        if init != 'bn' and tuple(w[name].shape) != tuple(shape):
            raise ValueError('%s has shape %s, expected %s' % (name, w[name].shape, shape))
    return w, num_objs, num_frags
