# coding: utf-8
"""GPU parity of the WaveNet training step (libwn_train_b200.so through its C ABI) against oracle/train_oracle.py.
fp32 mode must agree to 1e-4 (north_star tolerance on float outputs); bf16 mode -- the performance configuration of
BASELINE configs[3] -- is checked for closeness of loss and gradient direction, and by an overfitting property."""
import numpy as np
import pytest
import torch

from tacotron_wavenet_vocoder_korean_b200 import synth
from tests.train_helpers import train_case, rel_err, cosine, well_conditioned, at_cell_centres

pytestmark = pytest.mark.gpu

HP = dict(wavenet_learning_rate=1e-3, wavenet_decay_rate=0.5, wavenet_decay_steps=300000, wavenet_clip_gradients=False)


def _trainer(kw, T, dtype, w):
    from tacotron_wavenet_vocoder_korean_b200.wavenet.train import WaveNetTrainer
    tr = WaveNetTrainer(T, dtype=dtype, **kw)
    tr.load_state_dict(w)
    return tr


def _oracle(kw, w, dtype=torch.float32):
    from oracle import train_oracle as to
    return to.TorchWaveNetTrain(w, dtype=dtype, **kw)


def _grad_outliers(g, kw, w, args, l2=None):
    """Gradient criterion.  The reference evaluates cdf_plus - cdf_min in fp32 (mixture.py:60); the cancellation leaves a
    few 1e-4..1e-3 of relative noise in ANY fp32 evaluation of its gradients (tests/test_train_oracle.py::
    test_fp32_gradient_noise_floor).  So each tensor is compared with the fp64 oracle and must be within
    max(5e-4, 3x the fp32 oracle's own deviation from fp64) relative L2."""
    _, g32 = _oracle(kw, w).loss_and_grads(*args, l2)
    _, g64 = _oracle(kw, w, torch.float64).loss_and_grads(*args, l2)
    bad = {}
    for k in g64:
        mine, floor = rel_err(g[k], g64[k]), rel_err(g32[k], g64[k])
        if mine > max(5e-4, 3 * floor) and np.abs(g[k] - g64[k]).max() > 1e-6:
            bad[k] = (mine, floor)
    return bad


VARIANTS = {
    'full': {},
    'no_lc': dict(local_condition_channels=None, upsample_factor=None),
    'no_gc': dict(global_condition_channels=None, global_condition_cardinality=None),
    'no_bias': dict(use_biases=False),
    'wide': dict(residual_channels=32, dilation_channels=16, skip_channels=32, initial_filter_width=32, dilations=[1, 2, 4, 8, 16]),
}


@pytest.mark.parametrize('variant', sorted(VARIANTS))
def test_fp32_loss_logits_and_every_gradient_match_oracle(variant):
    kw = dict(synth.tiny_train(3), **VARIANTS[variant])
    T = 96 if variant != 'wide' else 120
    w, wav, mel, gc = train_case(kw, T)
    l2 = 0.01 if variant == 'full' else None
    Lo, go = _oracle(kw, w).loss_and_grads(wav, mel, gc, l2)
    tr = _trainer(kw, T, 'fp32', w)
    L = float(tr.loss_and_grads(wav, mel, gc, l2).item())
    raw_o = _oracle(kw, w).raw_output(wav, mel, gc)[0].detach().numpy().reshape(-1, kw['out_channels'])
    raw = tr.debug_get('raw_output').reshape(-1, kw['out_channels'])
    assert np.abs(raw - raw_o).max() <= 1e-4                      # north_star: 1e-4 on MoL logits
    assert abs(L - Lo) <= 1e-4 * max(1.0, abs(Lo))
    g = tr.state_dict('grads')
    assert set(g) == set(go)
    bad = _grad_outliers(g, kw, w, (wav, mel, gc), l2)
    assert not bad, bad


ONE_HOT = {
    'mulaw': lambda: synth.tiny_mulaw(2),                                        # BASELINE configs[0] family: no conditioning
    'mulaw_lc_gc': lambda: dict(synth.tiny_train(3), scalar_input=False),
    'mulaw_q64_no_bias': lambda: dict(synth.tiny_mulaw(2), quantization_channels=64, use_biases=False),
}


@pytest.mark.parametrize('variant', sorted(ONE_HOT))
def test_fp32_one_hot_input_softmax_head_matches_oracle(variant):
    """scalar_input=False: mu-law one-hot input, 2-tap causal layer over Q channels, softmax cross-entropy (model.py:257-296)."""
    kw = ONE_HOT[variant]()
    Q = kw['quantization_channels']
    T = 96
    w, wav, mel, gc = train_case(kw, T)
    wav = at_cell_centres(wav, Q)
    l2 = 0.01 if variant == 'mulaw' else None
    Lo, go = _oracle(kw, w).loss_and_grads(wav, mel, gc, l2)
    tr = _trainer(kw, T, 'fp32', w)
    i = tr.info()
    assert i['receptive_field'] == sum(kw['dilations']) + 2 and i['output_width'] == T - i['receptive_field']
    L = float(tr.loss_and_grads(wav, mel, gc, l2).item())
    raw_o = _oracle(kw, w).raw_output(wav, mel, gc)[0].detach().numpy().reshape(-1, Q)
    raw = tr.debug_get('raw_output').reshape(-1, Q)
    assert np.abs(raw - raw_o).max() <= 1e-4
    assert abs(L - Lo) <= 1e-4 * max(1.0, abs(Lo))
    g = tr.state_dict('grads')
    assert set(g) == set(go) and g['wavenet/conv1d/kernel'].shape == (2, Q, kw['residual_channels'])
    bad = _grad_outliers(g, kw, w, (wav, mel, gc), l2)
    assert not bad, bad


def test_bf16_one_hot_model_on_the_tcgen05_path():
    """R = D = 128 with the softmax head: the fused layer kernels are the same, only the causal layer and the loss differ."""
    kw = dict(synth.cfg2(2), scalar_input=False)
    T = 3600
    w, wav, mel, gc = train_case(kw, T)
    wav = at_cell_centres(wav)
    Lo, go = _oracle(kw, w).loss_and_grads(wav, mel, gc)
    tr = _trainer(kw, T, 'bf16', w)
    L = float(tr.loss_and_grads(wav, mel, gc).item())
    assert tr.info()['fused_launches'] == 5 * 30
    assert abs(L - Lo) <= 1e-2 * abs(Lo), (L, Lo)
    g = tr.state_dict('grads')
    big = [k for k in go if np.linalg.norm(go[k]) > 1e-4]
    bad = {k: cosine(g[k], go[k]) for k in big if cosine(g[k], go[k]) < 0.99}
    assert not bad, bad
    losses = [float(tr.train_step(wav, mel, gc, dict(HP, wavenet_learning_rate=1e-4)).item()) for _ in range(8)]
    assert np.all(np.isfinite(losses)) and losses[-1] < losses[0], losses


def test_state_dict_round_trip_and_layout():
    kw = synth.tiny_train(2)
    w, *_ = train_case(kw, 48)
    tr = _trainer(kw, 48, 'fp32', w)
    got = tr.state_dict()
    assert set(got) == set(synth.weight_shapes(**kw)) and list(got) == tr.variable_names
    assert list(got)[4:8] == ['wavenet/dilated_stack/layer0/dilation_layer/conv_filter/' + x for x in ('kernel', 'bias')] + \
        ['wavenet/dilated_stack/layer0/dilation_layer/conv_gate/' + x for x in ('kernel', 'bias')]   # creation order (model.py:68-69)
    assert all(np.array_equal(got[k], w[k]) for k in w)
    ema = tr.state_dict('ema')
    assert all(np.array_equal(ema[k], w[k]) for k in w)          # shadows start at the initial values
    i = tr.info()
    assert i['n_trainable'] == sum(int(np.prod(v.shape)) for v in w.values())
    assert i['receptive_field'] == 2 * (1 + 2 + 4) + 1 + 7 and i['output_width'] == 48 - i['receptive_field']


def test_adam_decay_ema_steps_match_oracle():
    """apply_gradients + EMA (model.py:325-346) against the float64 numpy restatement, fed the SAME gradients (Adam's first
    steps move every weight by lr*sign(g), so independently computed gradients would turn 1e-4 gradient noise into 2e-3
    parameter differences wherever a gradient is near zero)."""
    from oracle import train_oracle as to
    kw = synth.tiny_train(2)
    T = 72
    w, wav, mel, gc = train_case(kw, T)
    tr = _trainer(kw, T, 'fp32', w)
    hp = dict(HP, wavenet_decay_steps=2)                           # make the decay visible within 3 steps
    params = {k: v.astype(np.float64) for k, v in w.items()}
    m = {k: np.zeros_like(v) for k, v in params.items()}
    v = {k: np.zeros_like(vv) for k, vv in params.items()}
    ema = {k: vv.copy() for k, vv in params.items()}
    for step in range(3):
        tr.loss_and_grads(wav, mel, gc)
        g = tr.state_dict('grads')
        to.adam_ema_step(params, g, m, v, ema, step + 1, to.learning_rate(1e-3, step, 2, 0.5))
        tr.apply(to.learning_rate(1e-3, tr.global_step, 2, 0.5))
    got, got_ema, got_m, got_v = tr.state_dict(), tr.state_dict('ema'), tr.state_dict('adam_m'), tr.state_dict('adam_v')
    for k in params:
        assert np.abs(got[k] - params[k]).max() <= 2e-6, k
        assert np.abs(got_ema[k] - ema[k]).max() <= 2e-6, k
        np.testing.assert_allclose(got_m[k], m[k], rtol=1e-4, atol=1e-9)
        np.testing.assert_allclose(got_v[k], v[k], rtol=1e-4, atol=1e-12)
    assert tr.global_step == 3
    # train_step = loss_and_grads + decayed-lr apply
    before = tr.state_dict()
    tr.train_step(wav, mel, gc, hp)
    assert tr.global_step == 4 and any(np.abs(tr.state_dict()[k] - before[k]).max() > 0 for k in before)


def test_clip_by_global_norm():
    kw = synth.tiny_train(2)
    T = 72
    w, wav, mel, gc = train_case(kw, T)
    w = {k: v * (4.0 if k.endswith('conv1d_2/kernel') else 1.0) for k, v in w.items()}    # make the norm exceed 1
    tr = _trainer(kw, T, 'fp32', w)
    tr.loss_and_grads(wav, mel, gc)
    g = tr.state_dict('grads')
    gn = np.sqrt(sum(float((x.astype(np.float64) ** 2).sum()) for x in g.values()))
    assert gn > 1.5
    tr.apply(1e-3, clip_norm=1.0)
    m = tr.state_dict('adam_m')
    for k in ('wavenet/conv1d_1/kernel', 'wavenet/gc_embedding', 'wavenet/conv1d_2/bias'):
        np.testing.assert_allclose(m[k], 0.1 * g[k] / gn, rtol=1e-4, atol=1e-9)


def test_fp32_reference_layer_sizes_match_oracle():
    """30 layers, R=D=128, S=512, 80-channel mel, hop 300 (the layer sizes of BASELINE configs[3]) on a short crop."""
    kw = synth.cfg2(2)
    T = 3600
    w, wav, mel, gc = train_case(kw, T)
    Lo, go = _oracle(kw, w).loss_and_grads(wav, mel, gc)
    tr = _trainer(kw, T, 'fp32', w)
    L = float(tr.loss_and_grads(wav, mel, gc).item())
    assert abs(L - Lo) <= 1e-4 * max(1.0, abs(Lo))
    g = tr.state_dict('grads')
    bad = _grad_outliers(g, kw, w, (wav, mel, gc))
    assert not bad, bad


def test_bf16_step_close_to_fp32_oracle():
    kw = synth.tiny_train(3)
    T = 96
    w, wav, mel, gc = train_case(kw, T)
    Lo, go = _oracle(kw, w).loss_and_grads(wav, mel, gc)
    tr = _trainer(kw, T, 'bf16', w)
    L = float(tr.loss_and_grads(wav, mel, gc).item())
    assert abs(L - Lo) <= 3e-2 * max(1.0, abs(Lo))
    g = tr.state_dict('grads')
    big = [k for k in go if np.linalg.norm(go[k]) > 1e-3]
    worst = min(cosine(g[k], go[k]) for k in big)
    assert worst >= 0.98, {k: cosine(g[k], go[k]) for k in big if cosine(g[k], go[k]) < 0.98}


def test_bf16_overfits_one_batch_at_reference_layer_sizes():
    """30-layer R=D=128 S=512 model (BASELINE configs[3] layers) on a short crop: the loss must fall steadily when the
    same batch is repeated, and stay finite."""
    kw = synth.cfg2(2)
    T = 3600
    w, wav, mel, gc = train_case(kw, T)
    tr = _trainer(kw, T, 'bf16', w)
    hp = dict(HP, wavenet_learning_rate=1e-4)
    losses = [float(tr.train_step(wav, mel, gc, hp).item()) for _ in range(12)]
    assert all(np.isfinite(losses))
    assert losses[-1] < losses[0] - 0.3, losses
    i = tr.info()
    assert i['gemm_launches'] > 0 and i['kernel_launches'] > 0 and i['flops_per_step'] > 0


def test_fused_tcgen05_path_matches_fp32_oracle_and_cublaslt_path(monkeypatch):
    """R = D = 128 (BASELINE configs[3] layer sizes): the tcgen05/TMEM/TMA kernels (persistent fused forward, gate backward,
    dx, MN-major split-K weight gradients) against the fp32 torch oracle and against the cuBLASLt path on the same inputs."""
    kw = synth.cfg2(2)
    T = 3600
    w, wav, mel, gc = train_case(kw, T)
    w = well_conditioned(w)
    Lo, go = _oracle(kw, w).loss_and_grads(wav, mel, gc)
    res = {}
    for mode in ('cublaslt', 'tcgen05'):
        if mode == 'cublaslt':
            monkeypatch.setenv('WNT_NO_FUSED', '1')
        else:
            monkeypatch.delenv('WNT_NO_FUSED', raising=False)
        tr = _trainer(kw, T, 'bf16', w)
        L = float(tr.loss_and_grads(wav, mel, gc).item())
        res[mode] = (L, tr.state_dict('grads'), tr.info())
        assert abs(L - Lo) <= 5e-3 * abs(Lo), (mode, L, Lo)
    assert res['cublaslt'][2]['fused_launches'] == 0
    assert res['tcgen05'][2]['fused_launches'] == 30 + 2 * 30 + 2 * 30          # forward, (gate, dx), (wfg, wlc+wd) per layer
    g_c, g_t = res['cublaslt'][1], res['tcgen05'][1]
    big = [k for k in go if np.linalg.norm(go[k]) > 1e-4]
    bad = {k: (cosine(g_t[k], go[k]), rel_err(g_t[k], go[k])) for k in big if cosine(g_t[k], go[k]) < 0.995 or rel_err(g_t[k], go[k]) > 0.1}
    assert not bad, bad
    assert min(cosine(g_t[k], g_c[k]) for k in big) >= 0.995
    # the fused path must not be noticeably noisier than the cuBLASLt path
    assert max(rel_err(g_t[k], go[k]) for k in big) <= 1.5 * max(rel_err(g_c[k], go[k]) for k in big) + 0.01


def test_wavenet_model_add_loss_add_optimizer_surface():
    """The reference's call sequence (train_vocoder.py:100-123,169) on the drop-in class."""
    from tacotron_wavenet_vocoder_korean_b200.wavenet import WaveNetModel
    from tacotron_wavenet_vocoder_korean_b200.hparams import hparams
    kw = synth.tiny_train(2)
    T = 72
    w, wav, mel, gc = train_case(kw, T)
    net = WaveNetModel(train_mode=True, **kw)
    net.load_state_dict(w)
    net.add_loss(input_batch=torch.from_numpy(wav)[:, :, None], local_condition=mel, global_condition_batch=gc,
                 l2_regularization_strength=None, dtype='fp32')
    first = float(net.loss.item())
    Lo, _ = _oracle(kw, w).loss_and_grads(wav, mel, gc)
    assert abs(first - Lo) <= 1e-4 * max(1.0, abs(Lo))
    net.add_optimizer(hparams, global_step=0)
    for _ in range(5):
        net.optimize()
        net.add_loss(wav, mel, gc, dtype='fp32')
    assert float(net.loss.item()) < first
    assert abs(net.learning_rate - 1e-3 * 0.5 ** (4 / 300000.0)) < 1e-9
