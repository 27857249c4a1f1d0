# coding: utf-8
"""The oracles, the host code and the CUDA path against golden vectors produced by THE REFERENCE'S OWN Python, executed
unmodified on numpy TensorFlow stand-ins (tests/golden/tf_numpy_shim.py, tf_contrib_shim.py; one generating script per fixture,
tests/golden/make_reference_*.py): wavenet/model.py + mixture.py + ops.py (incremental generation, add_loss for both heads,
add_optimizer's plumbing), generate.py main() end to end (free-running and --wav_seed priming), the tacotron package and
synthesizer.py load / synthesize end to end, the vocoder data feeder, utils/audio.py, text/korean.py, utils/__init__.py.
This pins everything the reference's Python decides (wiring, variable names/shapes, queue order, conditioning alignment,
sampling formulas, loss construction, seeds, trimming, file naming); TensorFlow's own kernels are restated in the stand-ins,
hence tolerances of a few float32 ulps on floating-point outputs; integer outputs (mu-law samples, token ids, feeder batches,
frame counts) are compared for equality."""
import os

import numpy as np
import pytest

import oracle
from oracle import train_oracle as to
from tacotron_wavenet_vocoder_korean_b200 import synth
from tests.helpers import make_inputs, oracle_model
from tests.train_helpers import train_case, at_cell_centres

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
MULAW_LC = dict(synth.tiny_mulaw(), local_condition_channels=20, upsample_factor=[2, 3], global_condition_channels=8,
                global_condition_cardinality=3)


def test_variable_names_shapes_and_queues_are_the_reference_s():
    g = np.load(os.path.join(GOLD, 'ref_mol.npz'))
    kw = synth.tiny_mol()
    assert sorted(g['variable_names'].tolist()) == sorted(synth.weight_shapes(**kw))
    q = dict(zip(g['queue_names'].tolist(), g['queue_shapes'].tolist()))
    assert q['wavenet/queue/causal_queue'] == str((2, kw['initial_filter_width'], 1))
    assert q['wavenet/queue/local_condition_queue'] == str((2, 2, kw['local_condition_channels']))
    # model.py:60 `'dilation_queue'.format(i)` has no placeholder: TF uniquifies the names
    assert [n for n in g['queue_names'].tolist() if 'dilation' in n] == ['wavenet/queue/dilation_queue'] + \
        ['wavenet/queue/dilation_queue_%d' % i for i in range(1, len(kw['dilations']))]
    assert [q[n] for n in g['queue_names'].tolist() if 'dilation' in n] == [str((2, d + 1, kw['residual_channels'])) for d in kw['dilations']]
    assert int(g['receptive_field']) == oracle.receptive_field(2, kw['dilations'], True, kw['initial_filter_width'])
    g2 = np.load(os.path.join(GOLD, 'ref_mulaw.npz'))
    assert sorted(g2['variable_names'].tolist()) == sorted(synth.weight_shapes(**MULAW_LC))


def test_oracle_matches_reference_mol_path():
    g = np.load(os.path.join(GOLD, 'ref_mol.npz'))
    kw = synth.tiny_mol()
    om = oracle_model(kw, synth.make_weights(**kw))
    T = g['outputs'].shape[1]
    inp = make_inputs(kw, T)
    lc = om.upsample(inp['mel'])
    np.testing.assert_allclose(lc, g['lc_up'], atol=2e-6)                                     # create_upsample (model.py:102-111)
    s, lg = om.generate(T, inp['forced_full'], inp['uniforms'], lc_up=lc, gc_ids=inp['gc_ids'], want_logits=True)
    assert np.abs(lg - g['raw_output']).max() < 2e-5                                          # north_star tolerance is 1e-4
    assert np.abs(s - g['outputs'][:, :, 0]).max() < 2e-5                                     # mixture.py:84-114 draw


def test_oracle_tensor_level_mol_draw_matches_reference_draws():
    """oracle.mol_sample (mixture.py:84-114 on a tensor of logits) fed the REFERENCE's own logits reproduces the reference's draws."""
    g = np.load(os.path.join(GOLD, 'ref_mol.npz'))
    kw = synth.tiny_mol()
    inp = make_inputs(kw, g['outputs'].shape[1])
    s = oracle.mol_sample(g['raw_output'], inp['uniforms'])
    assert np.abs(s - g['outputs'][:, :, 0]).max() < 2e-5


def test_oracle_matches_reference_mulaw_path():
    g = np.load(os.path.join(GOLD, 'ref_mulaw.npz'))
    kw = MULAW_LC
    om = oracle_model(kw, synth.make_weights(**kw))
    T = g['outputs'].shape[1]
    inp = make_inputs(kw, T)
    lc = om.upsample(inp['mel'])
    np.testing.assert_allclose(lc, g['lc_up'], atol=2e-6)
    _, lg = om.generate(T, inp['forced_full'], inp['uniforms'], lc_up=lc, gc_ids=inp['gc_ids'], want_logits=True)
    probs = np.stack([[oracle.softmax_probs(lg[n, t]) for t in range(T)] for n in range(lg.shape[0])])
    assert np.abs(probs - g['outputs']).max() < 2e-6                                          # model.py:243 float64 softmax


def test_training_loss_and_mu_law_match_reference_graph():
    g = np.load(os.path.join(GOLD, 'ref_train.npz'))
    kw = synth.tiny_train(3)
    w, wav, mel, gc = train_case(kw, 96)
    m = to.TorchWaveNetTrain(w, **kw)
    assert abs(float(m.loss(wav, mel, gc).detach()) - float(g['loss'])) < 2e-5 * abs(float(g['loss']))
    assert abs(float(m.loss(wav, mel, gc, 0.01).detach()) - float(g['loss_l2'])) < 2e-5 * abs(float(g['loss_l2']))
    enc = oracle.mu_law_encode(g['mu_grid'], 256)
    assert np.mean(enc != g['mu_encoded']) < 0.01 and np.abs(enc - g['mu_encoded']).max() <= 1   # cell-edge ties only
    np.testing.assert_allclose(oracle.mu_law_decode(np.arange(256, dtype=np.float32), 256), g['mu_decoded'], atol=1e-6)


def test_training_loss_matches_reference_graph_at_reference_layer_sizes():
    """add_loss of the reference at the BASELINE configs[3] layer sizes (30 layers, R = D = 128, S = 512), 2 crops x 3600 samples:
    the input of tests/test_train_gpu.py::test_fp32_reference_layer_sizes_match_oracle, so the CUDA step is tied to it too."""
    g = np.load(os.path.join(GOLD, 'ref_train_cfg2.npz'))
    kw = synth.cfg2(2)
    w, wav, mel, gc = train_case(kw, 3600)
    m = to.TorchWaveNetTrain(w, **kw)
    assert abs(float(m.loss(wav, mel, gc).detach()) - float(g['loss'])) < 2e-5 * abs(float(g['loss']))
    assert abs(float(m.loss(wav, mel, gc, 0.01).detach()) - float(g['loss_l2'])) < 2e-5 * abs(float(g['loss_l2']))


def test_optimizer_plumbing_matches_reference_add_optimizer():
    """tests/golden/make_reference_optimizer_golden.py recorded what the reference's add_optimizer (wavenet/model.py:314-346) asks
    TensorFlow for: exponential_decay(wavenet_learning_rate, global_step, wavenet_decay_steps, wavenet_decay_rate) with no staircase,
    AdamOptimizer(lr) with TF's default betas / epsilon, clip_by_global_norm(., 1.0) only behind wavenet_clip_gradients,
    ExponentialMovingAverage(0.9999) over every trainable variable, in the order gradients -> apply -> EMA."""
    import inspect
    import json
    from tacotron_wavenet_vocoder_korean_b200.hparams import hparams
    from tacotron_wavenet_vocoder_korean_b200.wavenet.train import WaveNetTrainer, learning_rate_at
    g = json.load(open(os.path.join(GOLD, 'ref_optimizer.json')))
    for k, v in g['hparams'].items():
        assert getattr(hparams, k) == v, k                                       # hparams.py mirror
    for clip in (True, False):
        r = g['clip_%s' % clip]
        assert r['calls'] == ['compute_gradients', 'apply_gradients', 'ema.apply']
        assert r['adam_args'] == ['lr'] and r['adam_kwargs'] == {} and r['ema_kwargs'] == {} and r['ema_over_all_trainables']
        assert r['exponential_decay']['extra_args'] == [] and r['exponential_decay']['extra_kwargs'] == {}      # staircase=False
        assert r['applied_clipped'] == clip and (r.get('clip_norm') == 1.0) == clip
    d = g['clip_True']['exponential_decay']
    for step in (0, 1, 1000, 300000, 450000):
        assert abs(learning_rate_at(hparams, step) - d['learning_rate'] * d['decay_rate'] ** (step / d['decay_steps'])) < 1e-15
    defaults = {k: v.default for k, v in inspect.signature(WaveNetTrainer.apply).parameters.items() if v.default is not inspect.Parameter.empty}
    assert (defaults['beta1'], defaults['beta2'], defaults['epsilon']) == (0.9, 0.999, 1e-8)      # tf.train.AdamOptimizer defaults
    assert defaults['ema_decay'] == g['clip_True']['ema_decay'] == 0.9999 and defaults['clip_norm'] == 0.0
    src = inspect.getsource(WaveNetTrainer.train_step)
    assert "clip_norm=1.0 if get('wavenet_clip_gradients'" in src                # the clip norm the reference passes, behind the same flag


ONE_HOT_CASES = {'': lambda: synth.tiny_mulaw(2), '_lc_gc': lambda: dict(synth.tiny_train(3), scalar_input=False)}


@pytest.mark.parametrize('suffix', sorted(ONE_HOT_CASES))
def test_one_hot_training_loss_matches_reference_graph(suffix):
    """add_loss with scalar_input=False: mu_law_encode -> _one_hot -> 2-tap causal layer -> softmax cross-entropy (model.py:257-296)."""
    g = np.load(os.path.join(GOLD, 'ref_train_onehot.npz'))
    kw = ONE_HOT_CASES[suffix]()
    w, wav, mel, gc = train_case(kw, 96)
    wav = at_cell_centres(wav, kw['quantization_channels'])
    m = to.TorchWaveNetTrain(w, **kw)
    for tag, l2 in (('loss', None), ('loss_l2', 0.01)):
        ref = float(g[tag + suffix])
        assert abs(float(m.loss(wav, mel, gc, l2).detach()) - ref) < 2e-5 * abs(ref)


@pytest.mark.parametrize('case,batch_size,gc_enable,dirs', [('two_speakers', 3, True, ('spk_a', 'spk_b')), ('one_speaker', 2, False, ('spk_a',))])
def test_crop_feeder_yields_the_batches_of_the_reference_feeder(tmp_path, case, batch_size, gc_enable, dirs):
    """tests/golden/make_reference_feeder_golden.py ran the reference's DataFeederWavenet.make_batches() twice on this data set:
    path filtering by train.txt, the offset-2 start, skipped missing files, reshuffles on wrap-around, hop-aligned crops, the
    shuffle of 32 batches' worth of examples and the speaker ids must come out identically, batch for batch."""
    from tacotron_wavenet_vocoder_korean_b200.train_vocoder import WavenetCropFeeder
    from tests.train_helpers import make_feeder_dataset, FEEDER_HP
    g = np.load(os.path.join(GOLD, 'ref_feeder.npz'))
    make_feeder_dataset(str(tmp_path))
    hp = type('HP', (), dict(FEEDER_HP))()
    f = WavenetCropFeeder([str(tmp_path / d) for d in dirs], batch_size, 50, hp, gc_enable=gc_enable, seed=123,
                          crop_rng=np.random.RandomState(77))          # the reference crops with the global RNG, seeded 77 there
    assert f.sample_size == int(g[case + '_sample_size']) == 96
    ref_paths = dict(x.split(':') for x in g[case + '_path_dict'].tolist())
    n = int(g[case + '_n_batches'])
    assert n == 64
    for b in range(n):
        wav, mel, ids = next(f)
        np.testing.assert_array_equal(wav, g[case + '_wav'][b, :, :, 0])
        np.testing.assert_array_equal(mel, g[case + '_mel'][b])
        if gc_enable:
            np.testing.assert_array_equal(ids, g[case + '_ids'][b])
        else:
            assert ids is None
    assert {os.path.basename(k): ','.join(v) for k, v in f.path_dict.items()} == ref_paths      # same files kept, same final order


def test_attention_trim_matches_reference_synthesizer():
    """tests/golden/make_reference_trim_golden.py ran the reference's plot_graph_and_save_audio (synthesizer.py:198-277) on 400
    alignment paths; the number of frames it keeps must equal synthesizer.attention_trim_index on the same paths."""
    from tacotron_wavenet_vocoder_korean_b200.synthesizer import attention_trim_index
    g = np.load(os.path.join(GOLD, 'ref_trim.npz'))
    T, r = int(g['t_dec']), int(g['reduction_factor'])
    assert len(set(g['kept'].tolist())) > 10                          # the cases spread over early / late / never-finished endings
    for p, L, kept in zip(g['paths'], g['lens'], g['kept']):
        n_in = int(max(L, p.max() + 1))
        al = np.zeros((n_in, T), np.float32)
        al[p, np.arange(T)] = 1.0
        assert min(attention_trim_index(al, int(L), r), T * r) == int(kept), (p.tolist(), int(L))


# ---- the reference's own generate.py main() (tests/golden/make_reference_generate_golden.py) ---------------------------------
def _generate_main_inputs(kw, g):
    om = oracle_model(kw, synth.make_weights(**kw))
    N = kw['batch_size']
    lc = om.upsample(np.tile(g['mel'][None], (N, 1, 1)))                     # generate.py:153-155: ONE mel tiled over the batch
    return om, lc, np.full(N, 1, np.int32), np.random.RandomState(int(g['numpy_seed']))


def test_oracle_reproduces_reference_generate_main_mol():
    """generate.py main() of the reference, scalar input / MoL head / input_type 'raw', 2 rows x 60 samples: silent seed + one
    random sample (generate.py:186-188), per-sample loop (:202-233), output slice (:240) and save_wav."""
    from tacotron_wavenet_vocoder_korean_b200 import audio
    g = np.load(os.path.join(GOLD, 'ref_generate_main.npz'))
    kw = synth.tiny_mol(2)
    om, lc, gc, rs = _generate_main_inputs(kw, g)
    x0 = (2 * rs.rand(2) - 1).reshape(2, 1)                                  # generate.py:188, numpy's global RNG
    T = g['mol_wave'].shape[1]
    out = om.generate(T, x0, g['mol_uniforms'], lc_up=lc, gc_ids=gc)
    assert np.abs(out - g['mol_wave']).max() < 1e-4                          # free-running for 60 steps
    import io
    from scipy.io import wavfile
    for n in range(2):
        buf = io.BytesIO()
        audio.save_wav(out[n].copy(), buf, 24000)
        buf.seek(0)
        assert np.abs(wavfile.read(buf)[1].astype(np.int32) - g['mol_pcm'][n]).max() <= 1


@pytest.mark.parametrize('tag,temperature', [('mulaw_t1', 1.0), ('mulaw_t07', 0.7)])
def test_oracle_reproduces_reference_generate_main_mulaw(tag, temperature):
    """generate.py main() of the reference, one-hot input / softmax head / input_type 'mulaw-quantize': seed 128 ... + randint
    (:190-192), temperature rescaling and np.random.choice per row and step (:213-231), mu_law_decode (:246-247): the integer
    samples must be IDENTICAL."""
    g = np.load(os.path.join(GOLD, 'ref_generate_main.npz'))
    kw = MULAW_LC
    om, lc, gc, rs = _generate_main_inputs(kw, g)
    Q = kw['quantization_channels']
    wave = g[tag + '_wave']
    T = wave.shape[1]
    x0 = rs.randint(Q, size=2).reshape(2, 1).astype(np.float32)              # generate.py:192
    u = np.array([[rs.random_sample() for _ in range(2)] for _ in range(T)]).T      # np.random.choice: one double per row, step-major
    ids = om.generate(T, x0, u, lc_up=lc, gc_ids=gc, temperature=temperature)
    table = oracle.mu_law_decode(np.arange(Q, dtype=np.float32), Q)
    ref_ids = np.abs(wave[:, :, None] - table[None, None, :]).argmin(-1)
    assert np.abs(table[ref_ids] - wave).max() < 1e-6
    assert np.array_equal(ids.astype(np.int64), ref_ids)


def test_oracle_reproduces_reference_generate_main_with_wav_seed():
    """--wav_seed (generate.py:168-182): the first receptive_field samples of the seed prime the queues with ZERO local condition,
    the last of them enters the main loop together with lc row 0.  In this repository's terms: forced = seed[:rf], n_forced = rf,
    lc_shift = rf - 1, the first rf - 1 outputs are discarded."""
    g = np.load(os.path.join(GOLD, 'ref_generate_main.npz'))
    seed = g['seed_audio']
    # scalar input: the priming steps draw (and discard) samples too, so they consume uniforms
    kw = synth.tiny_mol(2)
    om, lc, gc, _ = _generate_main_inputs(kw, g)
    rf = oracle.receptive_field(2, kw['dilations'], True, kw['initial_filter_width'])
    T = g['mol_seeded_wave'].shape[1]
    forced = np.tile(seed[:rf][None], (2, 1))
    out = om.generate(rf - 1 + T, forced, g['mol_seeded_uniforms'], lc_up=lc, lc_shift=rf - 1, gc_ids=gc)
    assert np.abs(out[:, rf - 1:] - g['mol_seeded_wave']).max() < 1e-4
    # one-hot input: mu_law_encode of the seed; np.random.choice is first called in the main loop
    kw = MULAW_LC
    Q = kw['quantization_channels']
    om, lc, gc, rs = _generate_main_inputs(kw, g)
    rf = oracle.receptive_field(2, kw['dilations'], False, kw['initial_filter_width'])
    forced = np.tile(oracle.mu_law_encode(seed[:rf], Q).astype(np.float32)[None], (2, 1))
    u = np.zeros((2, rf - 1 + T))
    u[:, rf - 1:] = np.array([[rs.random_sample() for _ in range(2)] for _ in range(T)]).T
    ids = om.generate(rf - 1 + T, forced, u, lc_up=lc, lc_shift=rf - 1, gc_ids=gc)[:, rf - 1:]
    table = oracle.mu_law_decode(np.arange(Q, dtype=np.float32), Q)
    wave = g['mulaw_seeded_wave']
    assert np.array_equal(ids.astype(np.int64), np.abs(wave[:, :, None] - table[None, None, :]).argmin(-1))


# ---- the CUDA path against the same reference-generated vectors ------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize('name', ['ref_mol', 'ref_mulaw'])
def test_cuda_generation_matches_reference_goldens(name):
    import torch
    from tacotron_wavenet_vocoder_korean_b200.wavenet import WaveNetModel
    g = np.load(os.path.join(GOLD, name + '.npz'))
    kw = synth.tiny_mol() if name == 'ref_mol' else MULAW_LC
    net = WaveNetModel(train_mode=False, **kw)
    net.load_state_dict(synth.make_weights(**kw))
    T = g['outputs'].shape[1]
    inp = make_inputs(kw, T)
    lc = net.create_upsample(inp['mel'])
    np.testing.assert_allclose(lc.cpu().numpy(), g['lc_up'], atol=2e-6)
    s, lg = net.generate(T, inp['forced_full'], inp['uniforms'], lc_up=lc, gc_ids=inp['gc_ids'], want_logits=True)
    s, lg = s.cpu().numpy(), lg.cpu().numpy()
    if name == 'ref_mol':
        assert np.abs(lg - g['raw_output']).max() < 1e-4 and np.abs(s - g['outputs'][:, :, 0]).max() < 1e-4
    else:
        p = torch.softmax(torch.from_numpy(lg).double(), -1).float().numpy()
        assert np.abs(p - g['outputs']).max() < 1e-5


@pytest.mark.gpu
def test_cuda_path_reproduces_reference_generate_main():
    """The persistent kernel against the waveforms the reference's own generate.py main() produced (ref_generate_main.npz)."""
    from tacotron_wavenet_vocoder_korean_b200.wavenet import WaveNetModel
    g = np.load(os.path.join(GOLD, 'ref_generate_main.npz'))
    seed = int(g['numpy_seed'])
    for kw, tag, temperature in ((synth.tiny_mol(2), 'mol', 1.0), (MULAW_LC, 'mulaw_t1', 1.0), (MULAW_LC, 'mulaw_t07', 0.7)):
        net = WaveNetModel(train_mode=False, **kw)
        net.load_state_dict(synth.make_weights(**kw))
        lc = net.create_upsample(np.tile(g['mel'][None], (2, 1, 1)))
        rs = np.random.RandomState(seed)
        wave = g[tag + '_wave']
        T = wave.shape[1]
        if kw['scalar_input']:
            x0 = (2 * rs.rand(2) - 1).reshape(2, 1).astype(np.float32)
            out = net.generate(T, x0, g['mol_uniforms'], lc_up=lc, gc_ids=[1, 1]).cpu().numpy()
            assert np.abs(out - wave).max() < 1e-4
        else:
            Q = kw['quantization_channels']
            x0 = rs.randint(Q, size=2).reshape(2, 1).astype(np.float32)
            u = np.array([[rs.random_sample() for _ in range(2)] for _ in range(T)]).T
            ids = net.generate(T, x0, u, lc_up=lc, gc_ids=[1, 1], temperature=temperature).cpu().numpy()
            table = oracle.mu_law_decode(np.arange(Q, dtype=np.float32), Q)
            assert np.array_equal(ids.astype(np.int64), np.abs(wave[:, :, None] - table[None, None, :]).argmin(-1))


@pytest.mark.gpu
def test_cuda_training_loss_matches_reference_graph():
    from tacotron_wavenet_vocoder_korean_b200.wavenet.train import WaveNetTrainer
    g = np.load(os.path.join(GOLD, 'ref_train.npz'))
    kw = synth.tiny_train(3)
    w, wav, mel, gc = train_case(kw, 96)
    tr = WaveNetTrainer(96, dtype='fp32', **kw)
    tr.load_state_dict(w)
    assert abs(float(tr.loss_and_grads(wav, mel, gc).item()) - float(g['loss'])) < 1e-4 * abs(float(g['loss']))
    assert abs(float(tr.loss_and_grads(wav, mel, gc, 0.01).item()) - float(g['loss_l2'])) < 1e-4 * abs(float(g['loss_l2']))


@pytest.mark.gpu
@pytest.mark.parametrize('suffix', sorted(ONE_HOT_CASES))
def test_cuda_one_hot_training_loss_matches_reference_graph(suffix):
    from tacotron_wavenet_vocoder_korean_b200.wavenet.train import WaveNetTrainer
    g = np.load(os.path.join(GOLD, 'ref_train_onehot.npz'))
    kw = ONE_HOT_CASES[suffix]()
    w, wav, mel, gc = train_case(kw, 96)
    wav = at_cell_centres(wav, kw['quantization_channels'])
    tr = WaveNetTrainer(96, dtype='fp32', **kw)
    tr.load_state_dict(w)
    for tag, l2 in (('loss', None), ('loss_l2', 0.01)):
        ref = float(g[tag + suffix])
        assert abs(float(tr.loss_and_grads(wav, mel, gc, l2).item()) - ref) < 1e-4 * abs(ref)


# ---- utils/audio.py + hparams.py of the reference (tests/golden/make_reference_audio_golden.py) -----------------------------
def test_hparams_mirror_the_reference_values():
    from tacotron_wavenet_vocoder_korean_b200.hparams import hparams
    g = np.load(os.path.join(GOLD, 'ref_audio.npz'))
    for keys, vals in ((g['hparams_keys'], g['hparams_values']), (g['wn_keys'], g['wn_values'])):
        for k, v in zip(keys.tolist(), vals.tolist()):
            if k == 'use_lws':
                assert v == 0.0
                continue
            assert float(getattr(hparams, k)) == v, k
    assert list(hparams.dilations) == g['dilations'].tolist() and list(hparams.upsample_factor) == g['upsample_factor'].tolist()


def test_mel_oracle_and_save_wav_match_reference_glue(tmp_path):
    from scipy.io import wavfile
    from oracle import mel_oracle
    from tacotron_wavenet_vocoder_korean_b200 import audio
    g = np.load(os.path.join(GOLD, 'ref_audio.npz'))
    for i in range(3):
        np.testing.assert_allclose(mel_oracle.melspectrogram(g['wav%d' % i]), g['mel%d' % i], atol=1e-5)
    hp = mel_oracle.DEFAULTS
    m = hp['max_abs_value']
    S = g['norm_in']
    np.testing.assert_allclose(np.clip(2 * m * ((S - hp['min_level_db']) / -hp['min_level_db']) - m, -m, m), g['norm_out'], atol=1e-12)
    np.testing.assert_allclose(mel_oracle.preemphasis(g['wav0'][:64].astype(np.float64), 0.97), g['preemph'], atol=1e-7)
    x = np.logspace(-7, 1, 33)
    np.testing.assert_allclose(20 * np.log10(np.maximum(np.exp(hp['min_level_db'] / 20 * np.log(10)), x)), g['amp_to_db'], atol=1e-10)
    for i in range(2):                                                     # utils/audio.py:14-17
        p = str(tmp_path / 'x.wav')
        audio.save_wav(g['save_in%d' % i], p, 24000)
        sr, data = wavfile.read(p)
        assert sr == 24000 and np.array_equal(data, g['save_out%d' % i])


@pytest.mark.gpu
def test_cuda_melspectrogram_matches_reference_glue():
    from tacotron_wavenet_vocoder_korean_b200 import audio
    from tacotron_wavenet_vocoder_korean_b200.hparams import hparams
    g = np.load(os.path.join(GOLD, 'ref_audio.npz'))
    for i in range(3):
        got = audio.melspectrogram(g['wav%d' % i], hparams).cpu().numpy()
        assert got.shape == g['mel%d' % i].shape and np.abs(got - g['mel%d' % i]).max() <= 1e-4


# ---- tacotron/modules.py of the reference: prenet + encoder / post CBHG (tests/golden/make_reference_taco_golden.py) ---------
def test_taco_oracle_prenet_and_cbhg_match_reference_modules():
    from oracle.taco_oracle import TacotronOracle
    g = np.load(os.path.join(GOLD, 'ref_taco_modules.npz'))
    hp = synth.taco_tiny()
    w = synth.make_taco_weights(hp, 2)
    assert set(g['variable_names'].tolist()) <= set(w)                       # every name the reference's code creates exists
    o = TacotronOracle(hp, w, 2)
    p = g['x_emb']
    for i, _ in enumerate(hp['enc_prenet_sizes']):
        p = np.maximum(o._dense(p, 'prenet/dense_%d' % (i + 1)), 0)
    assert np.abs(p - g['prenet']).max() < 1e-5
    enc = o.cbhg(p, g['lengths'], 'encoder_cbhg', hp['enc_bank_size'], hp['enc_proj_sizes'], hp['enc_highway_depth'], hp['enc_rnn_size'],
                 g['before_highway'], g['rnn_init'])
    assert np.abs(enc - g['encoder_out']).max() < 2e-5
    post = o.cbhg(g['mel'], None, 'post_cbhg', hp['post_bank_size'], hp['post_proj_sizes'], hp['post_highway_depth'], hp['post_rnn_size'])
    assert np.abs(post - g['post_out']).max() < 2e-5


# ---- the WHOLE reference Tacotron graph: tacotron.py + rnn_wrappers.py + helpers.py + modules.py (make_reference_taco_full_golden.py) ----
TACO_FULL = ['tiny_mon_norm', 'tiny_mon', 'tiny_loc_sen', 'tiny_single_speaker', 'tiny_post_dense', 'tiny_mon_norm_manual',
             'full_mon_norm', 'full_loc_sen']      # full_*: the reference's hparams.py layer sizes (7.07 M parameters)


def _taco_case(tag):
    from tests.taco_helpers import case
    g = np.load(os.path.join(GOLD, 'ref_taco_full_%s.npz' % tag))
    hp, ns, w, ids, lens, spk, steps = case(tag.replace('_manual', ''))
    assert np.array_equal(ids, g['ids']) and np.array_equal(lens, g['lengths']) and int(g['steps']) == steps
    return g, hp, ns, w, ids, lens, spk, steps, (g['manual_alignments'] if 'manual_alignments' in g.files else None)


def test_host_side_of_synthesize_matches_reference_synthesizer():
    """tests/golden/make_reference_synth_golden.py ran the reference's Synthesizer.load + Synthesizer.synthesize unmodified on three
    Korean sentences for two speakers.  This repository's host steps (tokeniser, padding, input_lengths, attention trimming, output
    naming) around the Tacotron oracle must reproduce the sequences it fed, the network outputs and the trimmed mel files it wrote."""
    from oracle.taco_oracle import TacotronOracle
    from tacotron_wavenet_vocoder_korean_b200.synthesizer import attention_trim_index
    from tacotron_wavenet_vocoder_korean_b200.text import text_to_sequence, prepare_inputs
    g = np.load(os.path.join(GOLD, 'ref_synth_main.npz'))
    texts, spk = g['texts'].tolist(), g['speakers']
    sequences = prepare_inputs([text_to_sequence(t) for t in texts])                  # synthesizer.py:94-95
    assert np.array_equal(sequences, g['sequences'])
    lengths = np.array([int(np.argmax(a == 1)) + 1 for a in sequences])               # :126
    assert np.array_equal(lengths, g['input_lengths'])
    hp = synth.taco_tiny()
    steps = g['alignments'].shape[2]
    mel, lin, al = TacotronOracle(hp, synth.make_taco_weights(hp, 2, seed=4321), 2).synthesize(sequences, lengths, spk, max_iters=steps)
    assert np.abs(mel - g['mel_outputs']).max() < 2e-6 and np.abs(al - g['alignments']).max() < 1e-6
    assert g['files'].tolist() == ['s.%d.npy' % i for i in range(len(texts))]         # add_postfix(path, idx) -> <root>.<idx>.<ext>
    for i in range(len(texts)):
        end = attention_trim_index(al[i], len(sequences[i]), hp['reduction_factor'])   # :235-256
        ref = g['mel%d' % i]
        assert mel[i][:end].shape == ref.shape and np.abs(mel[i][:end] - ref).max() < 2e-6


@pytest.mark.parametrize('tag', TACO_FULL)
def test_taco_oracle_matches_reference_full_graph(tag):
    """Tacotron.initialize (inference) of the reference, run unmodified on the numpy TF stand-ins: the reference's AttentionWrapper
    with the manual-alignment override, DecoderPrenetWrapper, ConcatOutputAndAttentionWrapper, LocationSensitiveAttention,
    TacoTestHelper, cell stack, post net.  tf.contrib's cells / attention mechanisms / decode loop are restated in the stand-in."""
    from oracle.taco_oracle import TacotronOracle
    g, hp, ns, w, ids, lens, spk, steps, man = _taco_case(tag)
    mel, lin, al = TacotronOracle(hp, w, ns).synthesize(ids, lens, spk, max_iters=steps, manual_alignments=man)
    assert mel.shape == g['mel_outputs'].shape == (len(lens), steps * hp['reduction_factor'], hp['num_mels'])
    k = g['linear_outputs'].shape[-1]                                   # full-size fixtures keep the first 64 linear bins
    assert np.abs(mel - g['mel_outputs']).max() < 2e-6 and np.abs(lin[..., :k] - g['linear_outputs']).max() < 2e-6
    assert np.abs(al - g['alignments']).max() < 1e-6
    # every weight of this repository's state dict is a variable the reference graph creates (through the documented name map)
    assert set(g['mapped_names'].tolist()) == set(w)


@pytest.mark.gpu
@pytest.mark.parametrize('tag', TACO_FULL)
def test_cuda_tacotron_matches_reference_full_graph(tag):
    from tacotron_wavenet_vocoder_korean_b200.tacotron import Tacotron
    from tests.taco_helpers import Bag
    g, hp, ns, w, ids, lens, spk, steps, man = _taco_case(tag)
    m = Tacotron(Bag(hp))
    m.load_state_dict(w)
    if man is not None:
        m.is_manual_attention, m.manual_alignments = True, man
    m.initialize(ids, lens, ns, spk, rnn_decoder_test_mode=True, n_steps=steps)
    assert np.abs(m.mel_outputs.cpu().numpy() - g['mel_outputs']).max() <= 1e-4          # north_star tolerance on float mel
    assert np.abs(m.linear_outputs.cpu().numpy()[..., :g['linear_outputs'].shape[-1]] - g['linear_outputs']).max() <= 1e-4
    assert np.abs(m.alignments.cpu().numpy() - g['alignments']).max() <= 1e-4


# ---- text/korean.py of the reference (tests/golden/make_reference_text_golden.py) ------------------------------------------------
def test_korean_tokeniser_matches_reference_on_its_shipped_transcripts():
    """normalize + tokenize of the reference's own text/korean.py on the 160 transcripts it ships and on number / unit / letter
    cases (jamo package restated by Unicode arithmetic).  Without text/ko_dictionary.py's replacement tables installed, the
    tokeniser here must equal the reference run with those tables emptied; sentences the reference itself rejects are skipped."""
    import json
    from tacotron_wavenet_vocoder_korean_b200.text import korean as mk, text_to_sequence
    g = json.load(open(os.path.join(GOLD, 'ref_text.json'), encoding='utf-8'))
    assert g['symbols']['all_symbols'] == mk.ALL_SYMBOLS and g['symbols']['pad'] == mk.PAD and g['symbols']['eos'] == mk.EOS
    cases = [c for c in g['cases'] if 'error' not in c]
    assert len(cases) >= 180
    same_as_full = 0
    for c in cases:
        ids = list(mk.tokenize(c['text'], as_id=True))
        assert ids == c['ids_without_dictionaries'], c['text']
        if '~' not in c['normalized'] and '_' not in c['normalized']:     # text/__init__.py:_should_keep_symbol drops PAD / EOS characters
            assert list(text_to_sequence(c['text'])) == ids
        same_as_full += ids == c['ids']
    assert same_as_full >= len(cases) - 4          # only the few sentences that hit a dictionary entry (e.g. 'TV') differ


# ---- the benchmark configuration (BASELINE configs[1] layer sizes) through the reference's own wavenet/model.py -----------------
def test_oracle_matches_reference_at_cfg2_sizes():
    g = np.load(os.path.join(GOLD, 'ref_cfg2.npz'))
    kw = synth.cfg2(2)
    om = oracle_model(kw, synth.make_weights(**kw))
    T = g['outputs'].shape[1]
    inp = make_inputs(kw, T)
    lc = om.upsample(inp['mel'])[:, :T]
    np.testing.assert_allclose(lc, g['lc_up'], atol=3e-6)
    s, lg = om.generate(T, inp['forced_full'], inp['uniforms'], lc_up=lc, gc_ids=inp['gc_ids'], want_logits=True)
    assert np.abs(lg - g['raw_output']).max() < 5e-5 and np.abs(s - g['outputs'][:, :, 0]).max() < 5e-5
    assert int(g['receptive_field']) == 3101 and len(g['queue_names']) == 32          # causal + lc + 30 dilation queues


@pytest.mark.gpu
def test_cuda_generation_matches_reference_at_cfg2_sizes():
    from tacotron_wavenet_vocoder_korean_b200.wavenet import WaveNetModel
    g = np.load(os.path.join(GOLD, 'ref_cfg2.npz'))
    kw = synth.cfg2(2)
    net = WaveNetModel(train_mode=False, **kw)
    net.load_state_dict(synth.make_weights(**kw))
    T = g['outputs'].shape[1]
    inp = make_inputs(kw, T)
    lc = net.create_upsample(inp['mel'])
    s, lg = net.generate(T, inp['forced_full'], inp['uniforms'], lc_up=lc, gc_ids=inp['gc_ids'], want_logits=True)
    assert np.abs(lg.cpu().numpy() - g['raw_output']).max() < 1e-4 and np.abs(s.cpu().numpy() - g['outputs'][:, :, 0]).max() < 1e-4


# ---- utils/__init__.py helpers on the drop-in boundary (tests/golden/make_reference_utils_golden.py) -----------------------------
def test_load_json_and_load_hparams_match_reference(tmp_path, capsys):
    import json
    from tacotron_wavenet_vocoder_korean_b200.hparams import HParams, load_json, load_hparams
    g = json.load(open(os.path.join(GOLD, 'ref_utils.json'), encoding='utf-8'))
    p = tmp_path / 'params.json'
    p.write_text(g['params_text'], encoding='euc-kr')
    assert load_json(str(p)) == g['load_json']                       # trailing commas, euc-kr (utils/__init__.py:173-185)
    hp = HParams(sample_rate=24000, num_mels=80, dilations=[1, 2], name='x', upsample_factor=[5, 5, 12], hop_size=300)
    load_hparams(hp, str(tmp_path))
    assert hp.values() == g['load_hparams']                          # known keys overridden, unknown skipped (:156-172)
