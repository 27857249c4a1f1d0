# coding: utf-8
"""TensorFlow "tensor bundle" (checkpoint V2) reader / writer without TensorFlow.

The reference stores and restores its models with `tf.train.Saver` (utils/__init__.py:62-90, generate.py:157-161,
synthesizer.py:66-70): `<dir>/model.ckpt-<step>.index` + `.data-00000-of-00001` next to `params.json`.  This module
reads those two files into a `{variable name: numpy array}` dict -- the state_dict format of WaveNetModel /
Tacotron.load_state_dict -- and writes them, so that models trained here restore in the reference.

Format (tensorflow/core/util/tensor_bundle, tensorflow/core/lib/io/{table,block,format}; restated from the
published LevelDB table format -- TensorFlow itself is not installable here, so the only pins are the CRC-32C
check values, the LevelDB magic and the self round trip, see tests/test_tf_bundle.py):

  .index  an SSTable: data blocks of prefix-compressed (shared, non_shared, value_len varint32; key delta; value)
          entries + uint32 restart offsets + uint32 n_restarts, each block followed by a 1-byte compression type
          (0 none, 1 snappy) and a masked CRC-32C; an index block mapping last-key -> BlockHandle(varint64 offset,
          size); a 48-byte footer (metaindex handle, index handle, padding, magic 0xdb4775248b80fb57).
          key ""   -> BundleHeaderProto {1: num_shards, 2: endianness, 3: VersionDef{1: producer}}
          key name -> BundleEntryProto  {1: dtype, 2: TensorShapeProto{2: Dim{1: size}}, 3: shard_id, 4: offset,
                                         5: size, 6: fixed32 masked crc32c of the bytes, 7: slices (partitioned)}
  .data-SSSSS-of-NNNNN  raw little-endian tensor bytes at [offset, offset+size).
"""
import glob
import os
import struct

import numpy as np

TABLE_MAGIC = 0xdb4775248b80fb57
_MASK_DELTA = 0xa282ead8

# tensorflow/core/framework/types.proto
DT_FLOAT, DT_DOUBLE, DT_INT32, DT_UINT8, DT_INT16, DT_INT8, DT_STRING, DT_INT64, DT_BOOL = 1, 2, 3, 4, 5, 6, 7, 9, 10
DT_BFLOAT16, DT_UINT16, DT_HALF, DT_UINT32, DT_UINT64 = 14, 17, 19, 22, 23
_NP_OF_DT = {DT_FLOAT: np.float32, DT_DOUBLE: np.float64, DT_INT32: np.int32, DT_UINT8: np.uint8, DT_INT16: np.int16,
             DT_INT8: np.int8, DT_INT64: np.int64, DT_BOOL: np.bool_, DT_UINT16: np.uint16, DT_HALF: np.float16,
             DT_UINT32: np.uint32, DT_UINT64: np.uint64}
_DT_OF_NP = {np.dtype(v): k for k, v in _NP_OF_DT.items()}


class BundleError(ValueError):
    pass


# ---- CRC-32C (Castagnoli), byte-table recurrence; masked as in tensorflow/core/lib/hash/crc32c.h --------
def _make_table():
    t = []
    for i in range(256):
        c = i
        for _ in range(8):
            c = (c >> 1) ^ 0x82F63B78 if c & 1 else c >> 1
        t.append(c)
    return t


_CRC_TABLE = _make_table()
_CRC_NP = np.array(_CRC_TABLE, dtype=np.uint32)


def crc32c(data, crc=0):
    """CRC-32C of a bytes-like object.  Large buffers go through a multi-lane numpy evaluation + CRC combine."""
    buf = np.frombuffer(bytes(data) if not isinstance(data, (bytes, bytearray, memoryview)) else data, dtype=np.uint8)
    c = crc ^ 0xFFFFFFFF
    n = len(buf)
    if n >= 1 << 16:
        c = _crc_blocks(buf, c)
        return c ^ 0xFFFFFFFF
    t = _CRC_TABLE
    for b in buf.tolist():
        c = t[(c ^ b) & 0xFF] ^ (c >> 8)
    return c ^ 0xFFFFFFFF


def _gf2_times(mat, vec):
    s, i = 0, 0
    while vec:
        if vec & 1:
            s ^= mat[i]
        vec >>= 1
        i += 1
    return s


def _gf2_square(mat):
    return [_gf2_times(mat, mat[i]) for i in range(32)]


def _zeros_operator(nbytes):
    """32x32 GF(2) matrix advancing a (pre-inverted) CRC register over `nbytes` zero bytes (zlib's crc32_combine)."""
    odd = [0x82F63B78] + [1 << (i - 1) for i in range(1, 32)]     # one zero bit
    even = _gf2_square(odd)    # 2 bits
    odd = _gf2_square(even)    # 4 bits
    op = None
    n = nbytes
    mat = odd
    while n:
        mat = _gf2_square(mat)     # 8 bits = 1 byte, then doubles
        if n & 1:
            op = mat if op is None else [_gf2_times(mat, op[i]) for i in range(32)]
        n >>= 1
    return op


def _crc_blocks(buf, c):
    """Split the buffer in 256..16384 lanes, run the byte recurrence on all lanes at once in numpy, then combine."""
    n = len(buf)
    lanes = int(min(16384, max(256, n // 2048)))
    per = n // lanes
    head = buf[:per * lanes].reshape(lanes, per)
    regs = np.zeros(lanes, dtype=np.uint32)
    for j in range(per):
        regs = _CRC_NP[(regs ^ head[:, j]) & 0xFF] ^ (regs >> np.uint32(8))
    op = _zeros_operator(per)
    acc = c
    for r in regs.tolist():
        acc = _gf2_times(op, acc) ^ r      # crc(A||B) register = shift(reg_A, |B|) ^ reg_B(started from 0)
    t = _CRC_TABLE
    for b in buf[per * lanes:].tolist():
        acc = t[(acc ^ b) & 0xFF] ^ (acc >> 8)
    return acc


def mask_crc(c):
    return (((c >> 15) | (c << 17)) + _MASK_DELTA) & 0xFFFFFFFF


def unmask_crc(m):
    r = (m - _MASK_DELTA) & 0xFFFFFFFF
    return ((r >> 17) | (r << 15)) & 0xFFFFFFFF


# ---- varints / minimal protobuf ---------------------------------------------------------------------------------------
def _get_varint(b, p):
    r, s = 0, 0
    while True:
        if p >= len(b):
            raise BundleError("truncated varint")
        x = b[p]
        p += 1
        r |= (x & 0x7F) << s
        if x < 0x80:
            return r, p
        s += 7
        if s > 63:
            raise BundleError("varint too long")


def _put_varint(v):
    out = bytearray()
    v &= (1 << 64) - 1
    while v >= 0x80:
        out.append((v & 0x7F) | 0x80)
        v >>= 7
    out.append(v)
    return bytes(out)


def _pb_parse(b):
    """-> list of (field, wire_type, value); value is int (varint / fixed) or bytes (length-delimited)."""
    p, out = 0, []
    while p < len(b):
        key, p = _get_varint(b, p)
        f, wt = key >> 3, key & 7
        if wt == 0:
            v, p = _get_varint(b, p)
        elif wt == 1:
            v = struct.unpack_from('<Q', b, p)[0]
            p += 8
        elif wt == 2:
            n, p = _get_varint(b, p)
            v = bytes(b[p:p + n])
            if len(v) != n:
                raise BundleError("truncated protobuf field")
            p += n
        elif wt == 5:
            v = struct.unpack_from('<I', b, p)[0]
            p += 4
        else:
            raise BundleError("unsupported protobuf wire type %d" % wt)
        out.append((f, wt, v))
    return out


def _pb_field(f, wt, v):
    key = _put_varint((f << 3) | wt)
    if wt == 0:
        return key + _put_varint(v)
    if wt == 2:
        return key + _put_varint(len(v)) + v
    if wt == 5:
        return key + struct.pack('<I', v)
    raise BundleError("wire type")


def _parse_shape(b):
    dims = []
    for f, wt, v in _pb_parse(b):
        if f == 2 and wt == 2:
            size = 0
            for f2, wt2, v2 in _pb_parse(v):
                if f2 == 1:
                    size = v2 - (1 << 64) if v2 >= 1 << 63 else v2
            dims.append(size)
        elif f == 3 and v:
            raise BundleError("tensor of unknown rank")
    return tuple(dims)


def _parse_entry(b):
    e = dict(dtype=0, shape=(), shard_id=0, offset=0, size=0, crc32c=None, slices=0)
    for f, wt, v in _pb_parse(b):
        if f == 1:
            e['dtype'] = v
        elif f == 2:
            e['shape'] = _parse_shape(v)
        elif f == 3:
            e['shard_id'] = v
        elif f == 4:
            e['offset'] = v
        elif f == 5:
            e['size'] = v
        elif f == 6:
            e['crc32c'] = v
        elif f == 7:
            e['slices'] += 1
    return e


# ---- snappy (raw format) decoder: TF's table writer may compress blocks --------------------------------------------------
def _snappy_uncompress(b):
    n, p = _get_varint(b, 0)
    out = bytearray()
    while p < len(b):
        tag = b[p]
        p += 1
        kind = tag & 3
        if kind == 0:
            ln = tag >> 2
            if ln >= 60:
                nb = ln - 59
                ln = int.from_bytes(b[p:p + nb], 'little')
                p += nb
            ln += 1
            out += b[p:p + ln]
            p += ln
            continue
        if kind == 1:
            ln = ((tag >> 2) & 7) + 4
            off = ((tag >> 5) << 8) | b[p]
            p += 1
        elif kind == 2:
            ln = (tag >> 2) + 1
            off = b[p] | (b[p + 1] << 8)
            p += 2
        else:
            ln = (tag >> 2) + 1
            off = int.from_bytes(b[p:p + 4], 'little')
            p += 4
        if off == 0 or off > len(out):
            raise BundleError("corrupt snappy block")
        for _ in range(ln):
            out.append(out[-off])
    if len(out) != n:
        raise BundleError("snappy length mismatch")
    return bytes(out)


# ---- SSTable ---------------------------------------------------------------------------------------------------------
def _read_block(buf, offset, size, verify):
    raw = buf[offset:offset + size]
    trailer = buf[offset + size:offset + size + 5]
    if len(raw) != size or len(trailer) != 5:
        raise BundleError("block handle points outside the index file")
    ctype = trailer[0]
    if verify:
        want = unmask_crc(struct.unpack('<I', trailer[1:5])[0])
        if crc32c(bytes(raw) + bytes(trailer[:1])) != want:
            raise BundleError("index block checksum mismatch")
    if ctype == 0:
        return bytes(raw)
    if ctype == 1:
        return _snappy_uncompress(bytes(raw))
    raise BundleError("unknown block compression type %d" % ctype)


def _block_entries(block):
    if len(block) < 4:
        raise BundleError("block too small")
    n_restarts = struct.unpack_from('<I', block, len(block) - 4)[0]
    end = len(block) - 4 - 4 * n_restarts
    if end < 0:
        raise BundleError("corrupt restart array")
    p, key = 0, b''
    while p < end:
        shared, p = _get_varint(block, p)
        non_shared, p = _get_varint(block, p)
        vlen, p = _get_varint(block, p)
        if shared > len(key):
            raise BundleError("corrupt key prefix")
        key = key[:shared] + block[p:p + non_shared]
        p += non_shared
        yield key, block[p:p + vlen]
        p += vlen


def _read_table(buf, verify=True):
    if len(buf) < 48:
        raise BundleError("index file shorter than a table footer")
    footer = buf[-48:]
    if struct.unpack('<Q', footer[40:])[0] != TABLE_MAGIC:
        raise BundleError("not a TensorFlow checkpoint index (bad table magic)")
    _, p = _get_varint(footer, 0)
    _, p = _get_varint(footer, p)
    ioff, p = _get_varint(footer, p)
    isz, p = _get_varint(footer, p)
    items = []
    for _, handle in _block_entries(_read_block(buf, ioff, isz, verify)):
        boff, q = _get_varint(handle, 0)
        bsz, q = _get_varint(handle, q)
        items.extend(_block_entries(_read_block(buf, boff, bsz, verify)))
    return items


class BundleReader(object):
    """tf.train.load_checkpoint(prefix) look-alike: .get_variable_to_shape_map(), .get_tensor(name), .has_tensor."""

    def __init__(self, prefix, verify=True):
        self.prefix = prefix
        self.verify = verify
        with open(prefix + '.index', 'rb') as f:
            buf = f.read()
        self.entries = {}
        self.num_shards = 1
        for key, val in _read_table(buf, verify):
            if key == b'':
                for f_, wt, v in _pb_parse(val):
                    if f_ == 1:
                        self.num_shards = v
                    elif f_ == 2 and v != 0:
                        raise BundleError("big-endian bundles are not supported")
            else:
                self.entries[key.decode('utf-8')] = _parse_entry(val)
        self._shards = {}

    def has_tensor(self, name):
        return name in self.entries

    def get_variable_to_shape_map(self):
        return {k: list(e['shape']) for k, e in self.entries.items()}

    def get_variable_to_dtype_map(self):
        return {k: _NP_OF_DT.get(e['dtype']) for k, e in self.entries.items()}

    def _shard(self, i):
        if i not in self._shards:
            path = "%s.data-%05d-of-%05d" % (self.prefix, i, self.num_shards)
            self._shards[i] = np.memmap(path, dtype=np.uint8, mode='r') if os.path.getsize(path) else np.zeros(0, np.uint8)
        return self._shards[i]

    def get_tensor(self, name):
        if name not in self.entries:
            raise KeyError("tensor %r not found in checkpoint %s" % (name, self.prefix))
        e = self.entries[name]
        if e['slices']:
            raise BundleError("%s is a partitioned variable (slices): not supported" % name)
        if e['dtype'] not in _NP_OF_DT:
            raise BundleError("%s has unsupported dtype %d" % (name, e['dtype']))
        dt = np.dtype(_NP_OF_DT[e['dtype']])
        count = int(np.prod(e['shape'], dtype=np.int64)) if e['shape'] else 1
        if count * dt.itemsize != e['size']:
            raise BundleError("%s: size %d does not match shape %s" % (name, e['size'], e['shape']))
        raw = self._shard(e['shard_id'])[e['offset']:e['offset'] + e['size']]
        if len(raw) != e['size']:
            raise BundleError("%s: data shard is truncated" % name)
        raw = bytes(raw)
        if self.verify and e['crc32c'] is not None and crc32c(raw) != unmask_crc(e['crc32c']):
            raise BundleError("%s: tensor checksum mismatch" % name)
        return np.frombuffer(raw, dtype=dt).reshape(e['shape']).copy()


# ---- writer -----------------------------------------------------------------------------------------------------------
def _build_block(items, restart_interval=16):
    out, restarts, last = bytearray(), [], b''
    for i, (k, v) in enumerate(items):
        shared = 0
        if i % restart_interval == 0:
            restarts.append(len(out))
        else:
            m = min(len(k), len(last))
            while shared < m and k[shared] == last[shared]:
                shared += 1
        out += _put_varint(shared) + _put_varint(len(k) - shared) + _put_varint(len(v)) + k[shared:] + v
        last = k
    if not restarts:
        restarts = [0]
    for r in restarts:
        out += struct.pack('<I', r)
    out += struct.pack('<I', len(restarts))
    return bytes(out)


def _emit_block(f, block):
    off = f.tell()
    f.write(block)
    f.write(b'\x00' + struct.pack('<I', mask_crc(crc32c(block + b'\x00'))))
    return off, len(block)


def write_bundle(prefix, tensors, block_size=4096):
    """`tf.train.Saver().save` look-alike: {name: array} -> prefix.index + prefix.data-00000-of-00001 (one shard,
    uncompressed blocks, keys in byte order)."""
    os.makedirs(os.path.dirname(os.path.abspath(prefix)), exist_ok=True)
    names = sorted(tensors, key=lambda s: s.encode('utf-8'))
    items = [(b'', _pb_field(1, 0, 1) + _pb_field(3, 2, _pb_field(1, 0, 1)))]     # num_shards=1, little endian, version.producer=1
    offset = 0
    with open(prefix + '.data-00000-of-00001', 'wb') as f:
        for n in names:
            a = np.asarray(tensors[n], order='C')
            if a.dtype not in _DT_OF_NP:
                raise BundleError("%s: dtype %s cannot be stored" % (n, a.dtype))
            raw = a.tobytes()
            f.write(raw)
            shape = b''.join(_pb_field(2, 2, _pb_field(1, 0, int(d))) for d in a.shape)
            e = _pb_field(1, 0, _DT_OF_NP[a.dtype]) + _pb_field(2, 2, shape)
            if offset:
                e += _pb_field(4, 0, offset)
            e += _pb_field(5, 0, len(raw)) + _pb_field(6, 5, mask_crc(crc32c(raw)))
            items.append((n.encode('utf-8'), e))
            offset += len(raw)
    with open(prefix + '.index', 'wb') as f:
        index, cur, cur_bytes = [], [], 0
        for kv in items:
            cur.append(kv)
            cur_bytes += len(kv[0]) + len(kv[1]) + 3
            if cur_bytes >= block_size:
                off, sz = _emit_block(f, _build_block(cur))
                index.append((cur[-1][0], _put_varint(off) + _put_varint(sz)))
                cur, cur_bytes = [], 0
        if cur:
            off, sz = _emit_block(f, _build_block(cur))
            index.append((cur[-1][0], _put_varint(off) + _put_varint(sz)))
        moff, msz = _emit_block(f, _build_block([]))
        ioff, isz = _emit_block(f, _build_block(index, restart_interval=1))
        footer = _put_varint(moff) + _put_varint(msz) + _put_varint(ioff) + _put_varint(isz)
        f.write(footer + b'\x00' * (40 - len(footer)) + struct.pack('<Q', TABLE_MAGIC))


# ---- the reference's checkpoint-directory conventions --------------------------------------------------------------------
def get_most_recent_checkpoint(checkpoint_dir, checkpoint_step=None):
    """synthesizer.py:289-299 / utils/__init__.py:181-190: prefix of the newest `*.ckpt-<step>` in the directory."""
    if checkpoint_step is None:
        paths = glob.glob(os.path.join(checkpoint_dir, "*.ckpt-*.index"))
        if not paths:
            return None
        steps = [int(os.path.basename(p).split('-')[1].split('.')[0]) for p in paths]
        checkpoint_step = max(steps)
    return os.path.join(checkpoint_dir, "model.ckpt-%d" % checkpoint_step)


def checkpoint_state(logdir):
    """tf.train.get_checkpoint_state(logdir).model_checkpoint_path (utils/__init__.py:78): parses the text-proto
    `checkpoint` file; falls back to the newest index file."""
    path = os.path.join(logdir, 'checkpoint')
    if os.path.exists(path):
        with open(path, encoding='utf-8') as f:
            for line in f:
                if line.startswith('model_checkpoint_path:'):
                    p = line.split(':', 1)[1].strip().strip('"')
                    cand = p if os.path.isabs(p) else os.path.join(logdir, p)
                    if os.path.exists(cand + '.index'):
                        return cand
                    # a checkpoint directory copied from another machine records an absolute path that no longer exists
                    local = os.path.join(logdir, os.path.basename(p))
                    if os.path.exists(local + '.index'):
                        return local
                    break
    return get_most_recent_checkpoint(logdir)


_SKIP_SUFFIXES = ('/Adam', '/Adam_1')
_SKIP_NAMES = ('global_step', 'beta1_power', 'beta2_power')


def load_variables(prefix, use_ema=False, skip_queues=True, verify=True):
    """saver.restore (generate.py:157-161, synthesizer.py:66-70) -> {TF variable name: array}.  Optimizer slots
    (`.../Adam`, `.../Adam_1`, `beta*_power`), `global_step` and -- as generate.py:157 does -- the fast-generation
    queue variables are dropped.  `use_ema=True` substitutes the `<name>/ExponentialMovingAverage` shadow
    (wavenet/model.py:30,346) for each variable that has one."""
    r = BundleReader(prefix, verify=verify)
    out = {}
    ema = '/ExponentialMovingAverage'
    for name in r.entries:
        base = name.rsplit('/', 1)[-1]
        if name.endswith(_SKIP_SUFFIXES) or base in _SKIP_NAMES or name.endswith(ema):
            continue
        if skip_queues and 'queue' in name:
            continue
        src = name + ema if use_ema and r.has_tensor(name + ema) else name
        out[name] = r.get_tensor(src)
    return out


def global_step_of(prefix):
    """utils/__init__.py:82: the step is parsed from the file name."""
    return int(os.path.basename(prefix).split('-')[-1])
