# coding: utf-8
"""CPU tests of the training-step oracle (oracle/train_oracle.py) and of the C ABI surface of libwn_train_b200.so
(SURVEY.md 8f next-3).  TensorFlow cannot run here, so the oracle is pinned against independent restatements."""
import ctypes as C
import os
import re

import numpy as np
import pytest
import torch

from oracle import np_oracle, train_oracle as to
from tacotron_wavenet_vocoder_korean_b200 import synth, _train_lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_mol_loss_matches_float64_numpy_in_every_branch():
    rs = np.random.RandomState(0)
    B, T, K = 2, 400, 10
    y_hat = rs.randn(B, T, 3 * K).astype(np.float32)
    y_hat[..., 2 * K:] = rs.uniform(-9, -1, (B, T, K))
    y_hat[0, :40, 2 * K:] = -40.0                      # below log_scale_min: clamp branch
    y = rs.uniform(-1, 1, (B, T, 1)).astype(np.float32)
    y[0, :20] = -1.0                                   # y < -0.999
    y[1, :20] = 1.0                                    # y > 0.999
    y_hat[1, 100:140, K:2 * K] = 30.0                  # far means: cdf_delta <= 1e-5 -> log-pdf branch
    got = to.mol_loss(torch.from_numpy(y_hat).double(), torch.from_numpy(y).double()).numpy()
    ref = to.mol_loss_np(y_hat, y)
    np.testing.assert_allclose(got, ref, rtol=1e-9, atol=1e-9)
    got32 = to.mol_loss(torch.from_numpy(y_hat), torch.from_numpy(y)).numpy()
    np.testing.assert_allclose(got32, ref, rtol=5e-3, atol=1e-3)   # fp32: cdf_plus - cdf_min cancels


def test_training_graph_equals_incremental_generation_under_teacher_forcing():
    """Without local conditioning (the training graph aligns lc per layer) the training logits at
    position j are the incremental graph's output after it has consumed samples [0, rf-1+j]."""
    kw = synth.tiny_train(2)
    kw.update(local_condition_channels=None, upsample_factor=None)
    w = synth.make_weights(**kw)
    rs = np.random.RandomState(5)
    T = 60
    wav = rs.uniform(-1, 1, (2, T)).astype(np.float32)
    gc = np.array([2, 0])
    m = to.TorchWaveNetTrain(w, **kw)
    raw, _ = m.raw_output(wav, None, gc)
    raw = raw.detach().numpy()
    inc = np_oracle.NumpyWaveNet(**kw)
    inc.set_weights(w)
    outs = [inc.step(wav[:, t], None, gc) for t in range(T - 1)]
    rf = m.rf
    assert raw.shape[1] == T - rf
    for j in range(raw.shape[1]):
        np.testing.assert_allclose(raw[:, j], outs[rf - 1 + j], atol=2e-5)


def test_receptive_field_and_output_width():
    kw = synth.cfg2(1)
    m = to.TorchWaveNetTrain({}, **kw)
    assert m.rf == 3101
    kw = synth.cfg_hparams_default(1)
    assert to.TorchWaveNetTrain({}, **kw).rf == 5147   # hparams.py:79


def test_adam_decay_ema_against_torch_and_closed_forms():
    rs = np.random.RandomState(1)
    p0 = rs.randn(50)
    params, m, v, ema = {'a': p0.copy()}, {'a': np.zeros(50)}, {'a': np.zeros(50)}, {'a': p0.copy()}
    tp = torch.tensor(p0, dtype=torch.float64, requires_grad=True)
    opt = torch.optim.Adam([tp], lr=1e-3, betas=(0.9, 0.999), eps=1e-8)
    for t in range(1, 6):
        g = rs.randn(50)
        to.adam_ema_step(params, {'a': g}, m, v, ema, t, 1e-3)
        tp.grad = torch.tensor(g)
        opt.step()
    # torch divides by (sqrt(v_hat)+eps); TF by (sqrt(v)+eps) with lr_t: identical up to eps placement
    np.testing.assert_allclose(params['a'], tp.detach().numpy(), rtol=0, atol=2e-8)
    assert abs(to.learning_rate(1e-3, 300000, 300000, 0.5) - 5e-4) < 1e-12
    assert abs(to.learning_rate(1e-3, 150000, 300000, 0.5) - 1e-3 * 0.5 ** 0.5) < 1e-12
    # EMA closed form after one step
    e = p0 - (1 - 0.9999) * (p0 - 0.0)
    ema2 = {'a': p0.copy()}
    to.adam_ema_step({'a': np.zeros(50)}, {'a': np.zeros(50)}, {'a': np.zeros(50)}, {'a': np.zeros(50)}, ema2, 1, 1e-3)
    np.testing.assert_allclose(ema2['a'], e, atol=1e-15)
    # clip_by_global_norm(1.)
    params = {'a': np.zeros(4), 'b': np.zeros(4)}
    g = {'a': np.full(4, 3.0), 'b': np.full(4, 4.0)}          # norm 10
    m, v, ema = ({k: np.zeros(4) for k in g} for _ in range(3))
    to.adam_ema_step(params, g, m, v, ema, 1, 1e-3, clip=True)
    np.testing.assert_allclose(m['a'], 0.1 * 0.3)


def test_loss_gradient_matches_finite_differences():
    kw = synth.tiny_train(2)
    w = synth.make_weights(**kw)
    rs = np.random.RandomState(9)
    hop = int(np.prod(kw['upsample_factor']))
    T = 48
    wav = rs.uniform(-1, 1, (2, T)).astype(np.float32)
    mel = rs.randn(2, T // hop, kw['local_condition_channels']).astype(np.float32)
    gc = np.array([1, 2])
    m = to.TorchWaveNetTrain(w, dtype=torch.float64, **kw)
    L, g = m.loss_and_grads(wav, mel, gc, 0.01)
    for name, idx in (('wavenet/upsample1/kernel', (1, 0, 0, 0)), ('wavenet/gc_embedding', (1, 3)),
                      ('wavenet/dilated_stack/layer2/dilation_layer/lc_gate/kernel', (0, 5, 7))):
        w2 = {k: v.copy() for k, v in w.items()}
        eps = 1e-5
        w2[name] = w2[name].astype(np.float64)
        w2[name][idx] += eps
        Lp = float(to.TorchWaveNetTrain(w2, dtype=torch.float64, **kw).loss(wav, mel, gc, 0.01).detach())
        w2[name][idx] -= 2 * eps
        Lm = float(to.TorchWaveNetTrain(w2, dtype=torch.float64, **kw).loss(wav, mel, gc, 0.01).detach())
        assert abs((Lp - Lm) / (2 * eps) - g[name][idx]) < 1e-6 * max(1.0, abs(g[name][idx]))
    # the last layer's residual 1x1 feeds nothing (model.py:147): only the L2 term reaches it
    k = 'wavenet/dilated_stack/layer5/dilation_layer/dense/kernel'
    np.testing.assert_allclose(g[k], 0.01 * w[k].astype(np.float64), rtol=1e-6)
    assert np.all(m.loss_and_grads(wav, mel, gc, None)[1][k] == 0)


def test_fp32_gradient_noise_floor():
    """Why the GPU gradient tolerance is 2e-3 and not 1e-4: cdf_plus - cdf_min (mixture.py:60) cancels in fp32, so the
    fp32 evaluation of the reference graph already differs from the fp64 one by a few 1e-4 relative per tensor."""
    from tests.train_helpers import train_case, rel_err
    kw = synth.tiny_train(3)
    w, wav, mel, gc = train_case(kw, 96)
    L32, g32 = to.TorchWaveNetTrain(w, **kw).loss_and_grads(wav, mel, gc)
    L64, g64 = to.TorchWaveNetTrain(w, dtype=torch.float64, **kw).loss_and_grads(wav, mel, gc)
    assert abs(L32 - L64) < 1e-4 * abs(L64)
    errs = [rel_err(g32[k], g64[k]) for k in g64 if np.linalg.norm(g64[k]) > 0]
    assert 5e-5 < max(errs) < 2e-3


# ---- C ABI surface (no GPU needed) -----------------------------------------------------------------------------------
def test_train_cabi_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, 'include', 'wn_train_b200.h')).read()
    declared = sorted(set(re.findall(r'\b(wnt_[a-z_]+)\s*\(', hdr)))
    assert sorted(_train_lib.EXPORTS) == declared
    L = C.CDLL(_train_lib.LIB_PATH)
    for name in declared:
        assert hasattr(L, name), name


def test_train_config_struct_matches_header_and_validation():
    hdr = open(os.path.join(ROOT, 'include', 'wn_train_b200.h')).read()
    body = hdr[hdr.index('typedef struct wnt_config {'):hdr.index('} wnt_config;')]
    body = re.sub(r'/\*.*?\*/', '', body, flags=re.S)
    names = re.findall(r'\b([a-z_]+)(?:\[[A-Z_]+\])?\s*[,;]', body)
    assert names == [n for n, _ in _train_lib.WntConfig._fields_]
    cfg = _train_lib.make_config(7500, 'bf16', **synth.cfg2(64))
    assert cfg.n_layers == 30 and cfg.sample_size == 7500 and cfg.n_upsample == 3 and cfg.dtype == 0
    with pytest.raises(NotImplementedError):
        _train_lib.make_config(100, 'bf16', **dict(synth.tiny_train(), filter_width=3))
    with pytest.raises(ValueError):
        _train_lib.make_config(100, 'fp16', **synth.tiny_train())
    # without a GPU the library refuses to create a handle instead of falling back
    if not torch.cuda.is_available():
        L = _train_lib.lib()
        h = C.c_void_p()
        assert L.wnt_create(C.byref(cfg), C.byref(h)) < 0
        assert b'CUDA' in L.wnt_last_error(None) or b'device' in L.wnt_last_error(None)
