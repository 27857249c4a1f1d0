# coding: utf-8
"""Generates tests/golden/ref_audio.npz by running THE REFERENCE'S OWN utils/audio.py (melspectrogram, linearspectrogram,
preemphasis, _amp_to_db, _normalize, _denormalize, save_wav) with the reference's own hparams.py values.  Third-party
pieces that cannot be installed here are stubbed: `librosa.stft` / `librosa.filters.mel` by oracle/mel_oracle.py's
restatement (pinned against torch.stft / torchaudio in tests/test_mel.py), `tensorflow` by tests/golden/tf_numpy_shim.py.
So this pins the reference's GLUE -- order of operations, constants, clipping flags, dB floor, int16 scaling -- not librosa.

    python tests/golden/make_reference_audio_golden.py        (build container only)
"""
import importlib.util
import os
import sys
import tempfile
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get('REFERENCE_ROOT', '/root/reference')
sys.path.insert(0, HERE)
sys.path.insert(1, ROOT)
import tf_numpy_shim as tf            # noqa: E402
from oracle import mel_oracle         # noqa: E402


class HParams(object):                # tf.contrib.training.HParams: an attribute bag
    def __init__(self, **kw):
        self.__dict__.update(kw)

    def values(self):
        return dict(self.__dict__)


def install_stubs():
    tf.install()
    tf.contrib.training = types.SimpleNamespace(HParams=HParams)
    for name in ('tensorflow.contrib', 'tensorflow.contrib.training', 'tensorflow.contrib.training.python',
                 'tensorflow.contrib.training.python.training', 'tensorflow.contrib.training.python.training.hparam'):
        m = types.ModuleType(name)
        m.HParams = HParams
        sys.modules[name] = m
    lib = types.ModuleType('librosa')
    lib.stft = lambda y, n_fft, hop_length, win_length: mel_oracle.stft(y, n_fft, hop_length, win_length)
    lib.filters = types.ModuleType('librosa.filters')
    lib.filters.mel = lambda sr, n_fft, n_mels=128: mel_oracle.mel_basis(sr, n_fft, n_mels)
    lib.core = types.SimpleNamespace(load=None)
    lib.effects = types.SimpleNamespace(trim=None)
    lib.output = types.SimpleNamespace(write_wav=None)
    sys.modules['librosa'] = lib
    sys.modules['librosa.filters'] = lib.filters


def load(path, name):
    spec = importlib.util.spec_from_file_location(name, path)
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def main():
    install_stubs()
    hp = load(os.path.join(REF, 'hparams.py'), 'ref_hparams').hparams
    audio = load(os.path.join(REF, 'utils', 'audio.py'), 'ref_audio')
    rs = np.random.RandomState(42)
    t = np.arange(24000) / 24000.0
    wavs = [0.5 * np.sin(2 * np.pi * 220 * t) + 0.3 * np.sin(2 * np.pi * 3300 * t + 1.0) + 0.02 * rs.randn(len(t)),
            0.9 * np.sin(2 * np.pi * (100 + 4000 * t) * t) * np.hanning(len(t)),
            0.001 * rs.randn(7000)]
    out = {}
    for i, w in enumerate(wavs):
        w = w.astype(np.float32)
        out['wav%d' % i] = w
        out['mel%d' % i] = np.asarray(audio.melspectrogram(w, hp), np.float32)           # (80, frames)
        out['lin%d' % i] = np.asarray(audio.linearspectrogram(w, hp), np.float32)[:, :8]  # first frames only (size)
    S = np.linspace(-130, 10, 57)
    out['norm_in'] = S
    out['norm_out'] = np.asarray(audio._normalize(S, hp))
    out['denorm_out'] = np.asarray(audio._denormalize(np.linspace(-5, 5, 41), hp))
    out['amp_to_db'] = np.asarray(audio._amp_to_db(np.logspace(-7, 1, 33), hp))
    out['preemph'] = np.asarray(audio.preemphasis(wavs[0][:64], hp.preemphasis, hp.preemphasize))
    with tempfile.TemporaryDirectory() as d:
        from scipy.io import wavfile
        for i, w in enumerate((wavs[0][:4000].astype(np.float32), np.zeros(100, np.float32) + 1e-4)):
            p = os.path.join(d, 'x.wav')
            audio.save_wav(w.copy(), p, hp.sample_rate)                                   # utils/audio.py:14-17
            sr, data = wavfile.read(p)
            out['save_in%d' % i] = w
            out['save_out%d' % i] = data
            assert sr == hp.sample_rate
    keys = ('sample_rate', 'fft_size', 'hop_size', 'win_size', 'num_mels', 'preemphasis', 'preemphasize', 'min_level_db', 'ref_level_db',
            'max_abs_value', 'symmetric_mels', 'allow_clipping_in_normalization', 'signal_normalization', 'use_lws')
    out['hparams_keys'] = np.array(keys)
    out['hparams_values'] = np.array([float(getattr(hp, k)) for k in keys])
    # the WaveNet / training hyper-parameters this package mirrors (hparams.py:54-94)
    wk = ('filter_width', 'residual_channels', 'dilation_channels', 'skip_channels', 'quantization_channels', 'out_channels', 'gc_channels',
          'initial_filter_width', 'sample_size', 'wavenet_batch_size', 'wavenet_learning_rate', 'wavenet_decay_rate', 'wavenet_decay_steps',
          'l2_regularization_strength')
    out['wn_keys'] = np.array(wk)
    out['wn_values'] = np.array([float(getattr(hp, k)) for k in wk])
    out['dilations'] = np.array(hp.dilations)
    out['upsample_factor'] = np.array(hp.upsample_factor)
    np.savez_compressed(os.path.join(HERE, 'ref_audio.npz'), **out)
    print({k: np.shape(v) for k, v in out.items()})


if __name__ == '__main__':
    main()
