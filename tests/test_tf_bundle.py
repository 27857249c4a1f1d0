# coding: utf-8
"""tf_bundle.py: TF checkpoint-V2 reader/writer without TensorFlow.  TensorFlow is not
installable here, so the format is pinned by published check values (CRC-32C, crc masking, LevelDB magic, snappy
vectors) and by reader/writer round trips, including a hand-assembled index with multiple blocks and prefix
compression that the writer did not produce."""
import os
import struct

import numpy as np
import pytest

from tacotron_wavenet_vocoder_korean_b200 import tf_bundle as tb


def test_crc32c_check_values():
    # RFC 3720 B.4 / the "check" value of CRC-32C
    assert tb.crc32c(b'123456789') == 0xE3069283
    assert tb.crc32c(b'\x00' * 32) == 0x8A9136AA
    assert tb.crc32c(b'\xff' * 32) == 0x62A8AB43
    assert tb.crc32c(bytes(range(32))) == 0x46DD794E
    # leveldb/util/crc32c_test.cc: Mask(Value("foo")) round trips and differs from the value
    c = tb.crc32c(b'foo')
    assert tb.mask_crc(c) != c and tb.unmask_crc(tb.mask_crc(c)) == c
    assert tb.unmask_crc(tb.unmask_crc(tb.mask_crc(tb.mask_crc(c)))) == c


def test_crc32c_block_path_equals_bytewise():
    rs = np.random.RandomState(0)
    for n in (1 << 16, (1 << 16) + 77, 300001):
        b = rs.randint(0, 256, n).astype(np.uint8).tobytes()
        c = 0xFFFFFFFF
        for x in b:
            c = tb._CRC_TABLE[(c ^ x) & 0xFF] ^ (c >> 8)
        assert tb.crc32c(b) == c ^ 0xFFFFFFFF


def test_snappy_vectors():
    # literal only: length 5, tag (5-1)<<2
    assert tb._snappy_uncompress(bytes([5, 4 << 2]) + b'hello') == b'hello'
    # literal 'ab' + copy-1 (len 6, offset 2) -> 'abababab'
    assert tb._snappy_uncompress(bytes([8, 1 << 2]) + b'ab' + bytes([((6 - 4) << 2) | 1, 2])) == b'abababab'
    # copy-2: literal 'xyz' + copy len 3 offset 3
    assert tb._snappy_uncompress(bytes([6, 2 << 2]) + b'xyz' + bytes([((3 - 1) << 2) | 2, 3, 0])) == b'xyzxyz'
    with pytest.raises(tb.BundleError):
        tb._snappy_uncompress(bytes([4, 0 << 2]) + b'a' + bytes([((4 - 4) << 2) | 1, 9]))


def _state():
    rs = np.random.RandomState(3)
    s = {'wavenet/conv1d/kernel': rs.randn(32, 1, 16).astype(np.float32),
         'wavenet/gc_embedding': rs.randn(2, 32).astype(np.float32),
         'wavenet/postprocessing/conv1d_1/bias': np.zeros(30, np.float32),
         'global_step': np.array(12345, np.int64),
         'optimizer/beta1_power': np.array(0.5, np.float32),
         'wavenet/conv1d/kernel/Adam': rs.randn(32, 1, 16).astype(np.float32),
         'wavenet/conv1d/kernel/Adam_1': rs.randn(32, 1, 16).astype(np.float32),
         'wavenet/conv1d/kernel/ExponentialMovingAverage': rs.randn(32, 1, 16).astype(np.float32),
         'wavenet/queue/causal_queue': np.ones((1, 32, 1), np.float32)}
    for l in range(40):   # enough keys for several index blocks
        s['wavenet/dilated_stack/layer%d/dilation_layer/conv_filter/kernel' % l] = rs.randn(2, 8, 8).astype(np.float32)
    return s


def test_round_trip_and_filters(tmp_path):
    s = _state()
    prefix = str(tmp_path / 'model.ckpt-12345')
    tb.write_bundle(prefix, s, block_size=512)
    with open(prefix + '.index', 'rb') as f:
        raw = f.read()
    assert struct.unpack('<Q', raw[-8:])[0] == 0xdb4775248b80fb57        # leveldb kTableMagicNumber
    r = tb.BundleReader(prefix)
    assert set(r.entries) == set(s)
    assert r.get_variable_to_shape_map()['wavenet/conv1d/kernel'] == [32, 1, 16]
    for k, v in s.items():
        got = r.get_tensor(k)
        assert got.dtype == v.dtype and got.shape == v.shape and np.array_equal(got, v)
    v = tb.load_variables(prefix)
    assert 'global_step' not in v and not any('Adam' in k or 'queue' in k or 'beta1' in k or 'Exponential' in k for k in v)
    assert np.array_equal(v['wavenet/conv1d/kernel'], s['wavenet/conv1d/kernel'])
    v = tb.load_variables(prefix, use_ema=True)
    assert np.array_equal(v['wavenet/conv1d/kernel'], s['wavenet/conv1d/kernel/ExponentialMovingAverage'])
    assert tb.global_step_of(prefix) == 12345
    with pytest.raises(KeyError):
        r.get_tensor('nope')


def test_checkpoint_directory_conventions(tmp_path):
    d = str(tmp_path)
    for step in (1000, 25000, 3000):
        tb.write_bundle(os.path.join(d, 'model.ckpt-%d' % step), {'a': np.full(3, step, np.float32)})
    assert tb.get_most_recent_checkpoint(d).endswith('model.ckpt-25000')          # synthesizer.py:289-299
    assert tb.get_most_recent_checkpoint(d, 3000).endswith('model.ckpt-3000')
    assert tb.checkpoint_state(d).endswith('model.ckpt-25000')
    with open(os.path.join(d, 'checkpoint'), 'w') as f:                           # tf.train.get_checkpoint_state
        f.write('model_checkpoint_path: "model.ckpt-3000"\nall_model_checkpoint_paths: "model.ckpt-1000"\n')
    assert tb.checkpoint_state(d) == os.path.join(d, 'model.ckpt-3000')
    assert tb.load_variables(tb.checkpoint_state(d))['a'][0] == 3000
    assert tb.get_most_recent_checkpoint(str(tmp_path / 'empty')) is None


def test_corruption_is_detected(tmp_path):
    prefix = str(tmp_path / 'model.ckpt-1')
    tb.write_bundle(prefix, {'w': np.arange(100, dtype=np.float32)})
    with open(prefix + '.data-00000-of-00001', 'r+b') as f:
        f.seek(17)
        f.write(b'\x55')
    with pytest.raises(tb.BundleError, match='checksum'):
        tb.BundleReader(prefix).get_tensor('w')
    assert tb.BundleReader(prefix, verify=False).get_tensor('w').shape == (100,)
    with open(prefix + '.index', 'r+b') as f:
        f.seek(3)
        f.write(b'\x99')
    with pytest.raises(tb.BundleError):
        tb.BundleReader(prefix)
    with open(prefix + '.index', 'wb') as f:
        f.write(b'x' * 100)
    with pytest.raises(tb.BundleError, match='magic'):
        tb.BundleReader(prefix)


def test_hand_assembled_index_with_prefix_compression_and_snappy_block(tmp_path):
    """An index the writer here would never produce: restart interval 2 with shared prefixes, one data block stored
    as a snappy (all-literal) stream, offset field present, entry fields in a different order."""
    prefix = str(tmp_path / 'model.ckpt-7')
    a = np.arange(6, dtype=np.float32).reshape(2, 3)
    b = np.arange(4, dtype=np.int32)
    with open(prefix + '.data-00000-of-00001', 'wb') as f:
        f.write(a.tobytes() + b.tobytes())

    def entry(arr, dt, off):
        shape = b''.join(tb._pb_field(2, 2, tb._pb_field(1, 0, d)) for d in arr.shape)
        return (tb._pb_field(5, 0, arr.nbytes) + tb._pb_field(4, 0, off) + tb._pb_field(2, 2, shape) + tb._pb_field(1, 0, dt)
                + tb._pb_field(6, 5, tb.mask_crc(tb.crc32c(arr.tobytes()))))
    hdr = tb._pb_field(1, 0, 1)
    blk1 = tb._build_block([(b'', hdr), (b'model/a', entry(a, tb.DT_FLOAT, 0))], restart_interval=2)
    blk2 = tb._build_block([(b'model/ab', entry(b, tb.DT_INT32, a.nbytes))], restart_interval=2)
    # snappy "compress" blk2 as literals of <= 60 bytes
    comp = bytearray(tb._put_varint(len(blk2)))
    for i in range(0, len(blk2), 60):
        chunk = blk2[i:i + 60]
        comp += bytes([(len(chunk) - 1) << 2]) + chunk
    comp = bytes(comp)
    with open(prefix + '.index', 'wb') as f:
        o1, s1 = tb._emit_block(f, blk1)
        o2 = f.tell()
        f.write(comp + b'\x01' + struct.pack('<I', tb.mask_crc(tb.crc32c(comp + b'\x01'))))
        mo, ms = tb._emit_block(f, tb._build_block([]))
        idx = tb._build_block([(b'model/a', tb._put_varint(o1) + tb._put_varint(s1)),
                               (b'model/b', tb._put_varint(o2) + tb._put_varint(len(comp)))], restart_interval=1)
        io, isz = tb._emit_block(f, idx)
        foot = tb._put_varint(mo) + tb._put_varint(ms) + tb._put_varint(io) + tb._put_varint(isz)
        f.write(foot + b'\x00' * (40 - len(foot)) + struct.pack('<Q', tb.TABLE_MAGIC))
    r = tb.BundleReader(prefix)
    assert np.array_equal(r.get_tensor('model/a'), a) and np.array_equal(r.get_tensor('model/ab'), b)


def test_generate_and_synthesizer_loaders_read_bundles(tmp_path):
    """generate.load_checkpoint / tacotron.load_weights pick a TF checkpoint over weights.npz (no GPU needed)."""
    from tacotron_wavenet_vocoder_korean_b200 import synth
    from tacotron_wavenet_vocoder_korean_b200.generate import load_checkpoint
    from tacotron_wavenet_vocoder_korean_b200.tacotron import get_most_recent_checkpoint, load_weights
    kw = synth.tiny_mol(batch_size=2)
    w = synth.make_weights(**kw)
    d = str(tmp_path / 'wn')
    tb.write_bundle(os.path.join(d, 'model.ckpt-10'), w)
    got = load_checkpoint(d)
    assert set(got) == set(w) and all(np.array_equal(got[k], w[k]) for k in w)
    hp = synth.taco_tiny()
    tw = synth.make_taco_weights(hp, 2)
    d2 = str(tmp_path / 'taco')
    tb.write_bundle(os.path.join(d2, 'model.ckpt-500'), tw)
    tb.write_bundle(os.path.join(d2, 'model.ckpt-20'), {k: v * 0 for k, v in tw.items()})
    p = get_most_recent_checkpoint(d2)
    assert p.endswith('model.ckpt-500')
    got = load_weights(p)
    assert set(got) == set(tw) and all(np.array_equal(got[k], tw[k]) for k in tw)
    assert get_most_recent_checkpoint(d2, 20).endswith('model.ckpt-20')
    with pytest.raises(FileNotFoundError):
        load_checkpoint(str(tmp_path / 'nothing'))
