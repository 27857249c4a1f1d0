# coding: utf-8
"""Generates tests/golden/ref_generate_main.npz by running THE REFERENCE'S OWN generate.py `main()` -- argument parsing,
load_hparams from params.json, model construction, create_upsample of the tiled mel, queue initialisation, the silent seed with one
random sample, the per-sample loop (`window = waveform[:, -1:]`, `sess.run(next_sample, feed_dict)`, the temperature rescaling +
np.random.choice draw or the mixture-of-logistics draw, np.concatenate), the final slice / mu-law decode and audio.save_wav --
unmodified, on the numpy TensorFlow stand-in.

A `tf.Session` stand-in gives the eager stand-in graph semantics by re-tracing: WaveNetModel.predict_proba_incremental and
mu_law_decode are wrapped so that the tensor they return remembers how it was built; `sess.run(tensor, feed_dict)` stores the fed
values in the placeholders and builds it again (one `graph_pass`, variables reused by name -- the same emulation as
make_reference_goldens.py, driven by the reference's own loop this time).  TF's random_uniform (mixture.py:103,110) pops seeded
arrays; numpy's global RNG (initial sample, np.random.choice) is seeded with np.random.seed.

    python tests/golden/make_reference_generate_golden.py        (build container only)
"""
import contextlib
import glob
import io
import json
import os
import sys
import tempfile
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get('REFERENCE_ROOT', '/root/reference')
sys.path.insert(0, HERE)
import make_reference_audio_golden as mra      # noqa: E402
import tf_numpy_shim as tf                     # noqa: E402

NUMPY_SEED, UNIFORM_SEED, MEL_SEED, T_MEL = 11, 12, 13, 10


class Placeholder(object):
    """Hashable (feed_dict key) and array-like (every stand-in op starts with np.asarray)."""

    def __init__(self, dtype, shape=None, name=None):
        self.dtype = np.dtype(dtype)
        self.value = np.zeros([d if d is not None else 1 for d in shape], self.dtype)     # traced once before anything is fed

    def __array__(self, dtype=None, copy=None):
        return np.asarray(self.value, dtype=dtype)


class Session(object):
    def run(self, fetch, feed_dict=None):
        for p, v in (feed_dict or {}).items():
            p.value = np.asarray(v).astype(p.dtype)
        if callable(fetch):
            return fetch()
        rebuild = getattr(fetch, '_rebuild', None)
        if rebuild is None:
            return np.array(fetch)
        with tf.graph_pass():
            return np.array(rebuild())


def remember(fn):
    def wrapped(*a, **k):
        out = tf._t(fn(*a, **k))
        out._rebuild = lambda: fn(*a, **k)
        return out
    return wrapped


def run_main(kw, input_type, temperature, uniforms=None, seed_audio=None):
    """-> (float waveforms handed to save_wav, int16 wav files written, the mel)."""
    sys.path.insert(1, ROOT)
    from tacotron_wavenet_vocoder_korean_b200 import synth
    tf.reset()
    tf.set_initial_values(synth.make_weights(**kw))
    N = kw['batch_size']
    if uniforms is not None:
        nr = uniforms.shape[2] - 1
        tf.push_uniforms(np.full((N, 1, nr), 0.5), np.full((N, 1), 0.5))                 # consumed by the trace at graph-build time
        for t in range(uniforms.shape[1]):
            tf.push_uniforms(uniforms[:, t, None, :nr], uniforms[:, t, nr:])
    import generate as ref_gen                   # the reference module, unmodified
    assert ref_gen.__file__.startswith(REF)
    ref_gen.WaveNetModel.predict_proba_incremental = remember(ORIG['ppi'])
    ref_gen.mu_law_decode = remember(ORIG['decode'])
    handed = []
    real_save = ref_gen.audio.save_wav
    if seed_audio is not None:                   # --wav_seed: librosa.load / librosa.effects.trim are third party; the file content is given
        ref_gen.librosa.load = lambda filename, sr=None, mono=True: (np.asarray(seed_audio, np.float32), sr)
        ref_gen.audio.trim_silence = lambda wav, hp: wav

    def save_wav(wav, path, sr):
        handed.append(np.array(wav, copy=True))
        real_save(wav, path, sr)
    ref_gen.audio.save_wav = save_wav
    hop = int(np.prod(kw['upsample_factor']))
    mel = np.clip(np.random.RandomState(MEL_SEED).randn(T_MEL, kw['local_condition_channels']) * 1.5, -4, 4).astype(np.float32)
    with tempfile.TemporaryDirectory() as d:
        params = dict(dilations=kw['dilations'], filter_width=2, residual_channels=kw['residual_channels'], dilation_channels=kw['dilation_channels'],
                      quantization_channels=kw['quantization_channels'], out_channels=kw.get('out_channels', 30), skip_channels=kw['skip_channels'],
                      use_biases=kw['use_biases'], scalar_input=kw['scalar_input'], initial_filter_width=kw['initial_filter_width'],
                      gc_channels=kw['global_condition_channels'], num_mels=kw['local_condition_channels'], upsample_factor=kw['upsample_factor'],
                      hop_size=hop, sample_rate=24000, input_type=input_type)
        with open(os.path.join(d, 'params.json'), 'w', encoding='euc-kr') as f:
            json.dump(params, f)
        np.save(os.path.join(d, 'mel.npy'), mel)
        argv = ['generate.py', d, '--mel', os.path.join(d, 'mel.npy'), '--batch_size', str(N), '--logdir', os.path.join(d, 'out'),
                '--temperature', str(temperature), '--gc_cardinality', str(kw['global_condition_cardinality']), '--gc_id', '1']
        if seed_audio is not None:
            argv += ['--wav_seed', os.path.join(d, 'seed.wav')]
        old = sys.argv
        sys.argv = argv
        np.random.seed(NUMPY_SEED)
        try:
            with contextlib.redirect_stdout(io.StringIO()):
                ref_gen.main()
        finally:
            sys.argv = old
            ref_gen.audio.save_wav = real_save
        from scipy.io import wavfile
        files = sorted(glob.glob(os.path.join(d, 'out', 'generate', '*', 'test-*.wav')))
        assert len(files) == N, files
        pcm = np.stack([wavfile.read(f)[1] for f in files])
    return np.stack(handed), pcm, mel


ORIG = {}


def main():
    sys.path.insert(1, ROOT)
    from tacotron_wavenet_vocoder_korean_b200 import synth
    mra.install_stubs()
    tf.device = lambda *_: contextlib.nullcontext()
    tf.Session = Session
    tf.placeholder = Placeholder
    tf.global_variables = lambda: list(tf.S.variables.values())
    tf.train.Saver = lambda var_list=None: types.SimpleNamespace(restore=lambda sess, path: None)
    tf.train.get_checkpoint_state = lambda logdir: types.SimpleNamespace(model_checkpoint_path=logdir + '/model.ckpt-1234')
    tf.size = lambda x: int(np.asarray(x).size)                 # create_seed (generate.py:101-103)
    tf.constant = lambda v, *a, **k: v
    tf.cond = lambda pred, a, b: a() if pred else b()
    sys.path.insert(0, REF)
    for m in ('utils', 'hparams', 'wavenet', 'generate'):
        sys.modules.pop(m, None)
    import generate as ref_gen
    ORIG['ppi'] = ref_gen.WaveNetModel.predict_proba_incremental
    ORIG['decode'] = ref_gen.mu_law_decode
    out = {}
    # scalar input, mixture-of-logistics head, input_type 'raw' (the reference's hparams.py defaults), 2 rows x 60 samples
    kw = synth.tiny_mol(2)
    u = np.random.RandomState(UNIFORM_SEED).uniform(1e-5, 1 - 1e-5, (2, T_MEL * 6, kw['out_channels'] // 3 + 1)).astype(np.float32)
    wav, pcm, mel = run_main(kw, 'raw', 1.0, uniforms=u)
    out.update(mol_wave=wav, mol_pcm=pcm, mol_uniforms=u, mel=mel)
    # one-hot input, softmax head, input_type 'mulaw-quantize', temperature 1 and 0.7
    kw = dict(synth.tiny_mulaw(2), local_condition_channels=20, upsample_factor=[2, 3], global_condition_channels=8, global_condition_cardinality=3)
    for tag, temp in (('mulaw_t1', 1.0), ('mulaw_t07', 0.7)):
        wav, pcm, mel2 = run_main(kw, 'mulaw-quantize', temp)
        assert np.array_equal(mel, mel2)
        out[tag + '_wave'] = wav
        out[tag + '_pcm'] = pcm
    # --wav_seed priming (generate.py:168-182): the first receptive_field samples of the seed are fed with zero local condition
    rs = np.random.RandomState(14)
    seed_audio = np.clip(0.4 * np.sin(np.arange(40) * 0.3) + 0.05 * rs.randn(40), -1, 1).astype(np.float32)
    kw = synth.tiny_mol(2)
    rf = 1 + sum(kw['dilations']) + kw['initial_filter_width'] - 1
    u = np.random.RandomState(UNIFORM_SEED + 1).uniform(1e-5, 1 - 1e-5, (2, rf - 1 + T_MEL * 6, kw['out_channels'] // 3 + 1)).astype(np.float32)
    wav, pcm, _ = run_main(kw, 'raw', 1.0, uniforms=u, seed_audio=seed_audio)
    out.update(seed_audio=seed_audio, mol_seeded_wave=wav, mol_seeded_uniforms=u)
    kw = dict(synth.tiny_mulaw(2), local_condition_channels=20, upsample_factor=[2, 3], global_condition_channels=8, global_condition_cardinality=3)
    wav, pcm, _ = run_main(kw, 'mulaw-quantize', 1.0, seed_audio=seed_audio)
    out['mulaw_seeded_wave'] = wav
    np.savez_compressed(os.path.join(HERE, 'ref_generate_main.npz'), numpy_seed=np.int64(NUMPY_SEED), **out)
    print({k: (v.shape, v.dtype) for k, v in out.items()})
    print(out['mol_wave'][:, :6], out['mulaw_t1_wave'][:, :6])


if __name__ == '__main__':
    main()
