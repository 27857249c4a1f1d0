# coding: utf-8
"""Generates tests/golden/ref_feeder.npz by running THE REFERENCE'S OWN datasets/datafeeder_wavenet.py (DataFeederWavenet:
get_path_dict, the per-speaker example loop with its offset-2 start and reshuffles, the hop-aligned random crop, the
32-batch shuffle of make_batches, _prepare_batch) on a small synthetic data directory (tests/train_helpers.make_feeder_dataset).
TensorFlow's placeholders / FIFOQueue / Session are stand-ins that only record what is enqueued; librosa is not touched.

The reference crops with the GLOBAL numpy RNG (datafeeder_wavenet.py:153) and shuffles with RandomState(123) (:65); the global
RNG is seeded here and the same stream is handed to WavenetCropFeeder(crop_rng=...) by tests/test_reference_pin.py.

    python tests/golden/make_reference_feeder_golden.py        (build container only)
"""
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get('REFERENCE_ROOT', '/root/reference')
sys.path.insert(0, HERE)
import make_reference_audio_golden as mra      # noqa: E402  (stubs for tensorflow.contrib HParams and librosa)
import tf_numpy_shim as tf                     # noqa: E402

CROP_SEED = 77
CASES = {'two_speakers': dict(batch_size=3, gc_enable=True, dirs=('spk_a', 'spk_b')),
         'one_speaker': dict(batch_size=2, gc_enable=False, dirs=('spk_a',))}


class _Tensor(object):
    def __init__(self, shape=None):
        self.shape = shape

    def set_shape(self, shape):
        self.shape = shape


class _Queue(object):
    def __init__(self, capacity, dtypes, name=None):
        self.n = len(dtypes)

    def enqueue(self, placeholders):
        return ('enqueue', tuple(placeholders))

    def dequeue(self):
        return tuple(_Tensor() for _ in range(self.n))


class _Session(object):
    def __init__(self):
        self.fed = []

    def run(self, op, feed_dict=None):
        assert op[0] == 'enqueue'
        self.fed.append([feed_dict[p] for p in op[1]])


def main():
    sys.path.insert(0, ROOT)
    from tests.train_helpers import make_feeder_dataset, FEEDER_HP
    mra.install_stubs()
    tf.placeholder = lambda dtype, shape=None, name=None: _Tensor(shape)
    tf.FIFOQueue = _Queue
    sys.path.insert(0, REF)                       # `from utils import audio`, `from hparams import hparams` of the reference
    for m in ('utils', 'hparams', 'datasets'):
        sys.modules.pop(m, None)
    from datasets import datafeeder_wavenet as ref      # the reference module, unmodified
    assert ref.__file__.startswith(REF)
    for k, v in FEEDER_HP.items():
        setattr(ref.hparams, k, v)
    out = {}
    cwd = os.getcwd()
    for case, c in CASES.items():
        with tempfile.TemporaryDirectory() as d:
            make_feeder_dataset(d)
            os.chdir(d)                                  # relative data_dirs: the golden must not depend on the temp path
            try:
                np.random.seed(CROP_SEED)
                ref.hparams.skip_path_filter = False
                f = ref.DataFeederWavenet(None, list(c['dirs']), c['batch_size'], receptive_field=50, gc_enable=c['gc_enable'])
                f.sess, f._step = _Session(), 0
                f.make_batches()
                f.make_batches()
            finally:
                os.chdir(cwd)
        fed = f.sess.fed
        out[case + '_n_batches'] = np.int64(len(fed))
        out[case + '_wav'] = np.stack([np.stack(b[0]) for b in fed if len(b[0]) == c['batch_size']]).astype(np.float32)
        out[case + '_mel'] = np.stack([np.stack(b[1]) for b in fed if len(b[0]) == c['batch_size']]).astype(np.float32)
        if c['gc_enable']:
            out[case + '_ids'] = np.stack([np.asarray(b[2]) for b in fed if len(b[0]) == c['batch_size']]).astype(np.int32)
        out[case + '_sample_size'] = np.int64(f.sample_size)
        out[case + '_path_dict'] = np.array(['%s:%s' % (k, ','.join(v)) for k, v in sorted(f.path_dict.items())])
        print(case, len(fed), out[case + '_wav'].shape, out[case + '_mel'].shape, f.sample_size)
    np.savez_compressed(os.path.join(HERE, 'ref_feeder.npz'), **out)


if __name__ == '__main__':
    main()
