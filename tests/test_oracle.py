"""CPU tests of the oracle (test infrastructure): known-answer values from the reference's comments,
the committed golden fixtures, and an independent numpy restatement.  PARITY UNPINNED: the reference
itself cannot run here (no TensorFlow 1.x), see oracle/wn_oracle.c."""
import os

import numpy as np
import pytest
import torch

import oracle
from oracle import np_oracle
from tests.helpers import make_inputs, oracle_model, plan_from_dict
from tacotron_wavenet_vocoder_korean_b200 import synth

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
CYCLE = [1, 2, 4, 8, 16, 32, 64, 128, 256, 512]


def test_receptive_field_kats():
    # generate.py:192 comment: 50 layers, one-hot input, filter_width 2 -> 5117
    assert oracle.receptive_field(2, CYCLE * 5, False, 32) == 5117
    # hparams.py:59-79 defaults (scalar input, initial_filter_width 32) -> 5147
    assert oracle.receptive_field(2, CYCLE * 5, True, 32) == 5147
    assert oracle.receptive_field(2, CYCLE, False, 32) == 1025       # BASELINE cfg-1
    assert oracle.receptive_field(2, CYCLE * 3, True, 32) == 3101    # BASELINE cfg-2
    for args in [(2, CYCLE * 5, False, 32), (2, [1, 2, 4], True, 8), (2, [3, 5], False, 32)]:
        assert oracle.receptive_field(*args) == np_oracle.calculate_receptive_field(*args)


def test_upsample_factor_product_equals_hop():
    # hparams.py:79: np.prod(upsample_factor) must equal hop_size (300)
    assert int(np.prod(synth.cfg2()['upsample_factor'])) == 300


def test_mu_law_codec():
    g = np.load(os.path.join(GOLD, 'codec.npz'))
    enc = oracle.mu_law_encode(g['grid'], 256)
    assert np.array_equal(enc, g['enc'])
    enc_np = np_oracle.mu_law_encode(g['grid'], 256)        # libm log1p: may differ by one code on a cell edge
    assert np.mean(enc != enc_np) < 2e-3 and np.abs(enc - enc_np).max() <= 1
    assert enc.min() == 0 and enc.max() == 255
    codes = np.arange(256, dtype=np.float32)
    dec = oracle.mu_law_decode(codes, 256, True)
    assert np.array_equal(dec, g['dec_q'])
    np.testing.assert_allclose(dec, np_oracle.mu_law_decode(codes, 256, True), atol=1e-6)
    assert dec[0] == -1.0 and abs(dec[255] - 1.0) < 1e-6 and np.all(np.diff(dec) > 0)
    # decode then encode is the identity on all 256 codes (ops.py:22-47 are inverse up to quantisation)
    assert np.array_equal(oracle.mu_law_encode(dec, 256), np.arange(256))
    # encode then decode stays within one quantisation cell
    x = np.linspace(-1, 1, 2001).astype(np.float32)
    y = oracle.mu_law_decode(oracle.mu_law_encode(x, 256).astype(np.float32), 256, True)
    assert np.max(np.abs(x - y)) < 0.03
    np.testing.assert_array_equal(oracle.mu_law_decode(np.linspace(-1, 1, 513).astype(np.float32), 256, False), g['dec_c'])


def test_pinned_math_close_to_libm():
    xs = np.concatenate([np.linspace(-80, 80, 4001), np.linspace(-1, 1, 2001)]).astype(np.float32)
    e = np.array([oracle.math_probe(0, v) for v in xs], np.float64)
    np.testing.assert_allclose(e, np.exp(xs.astype(np.float64)), rtol=4e-7)
    pos = np.concatenate([np.logspace(-37, 37, 3000), np.linspace(0.5, 2, 1500)]).astype(np.float32)
    l = np.array([oracle.math_probe(1, v) for v in pos], np.float64)
    np.testing.assert_allclose(l, np.log(pos.astype(np.float64)), rtol=3e-7, atol=2e-7)
    t = np.array([oracle.math_probe(2, v) for v in xs], np.float64)
    np.testing.assert_allclose(t, np.tanh(xs.astype(np.float64)), atol=2.5e-7)
    s = np.array([oracle.math_probe(3, v) for v in xs], np.float64)
    np.testing.assert_allclose(s, 1 / (1 + np.exp(-xs.astype(np.float64))), rtol=5e-7, atol=1e-37)
    assert oracle.math_probe(0, -200.0) == 0.0 and oracle.math_probe(1, 0.0) == -np.inf
    d = np.array([oracle.math_probe(5, v) for v in np.linspace(-50, 0, 500)], np.float64)
    np.testing.assert_allclose(d, np.exp(np.linspace(-50, 0, 500).astype(np.float32).astype(np.float64)), rtol=1e-6)


def test_upsample_matches_conv_transpose2d():
    # create_upsample (model.py:102-111): conv2d_transpose(kernel (F,2), strides (F,1), 'same') is torch's
    # conv_transpose2d(stride=(F,1)) cropped to the first C columns
    kw = synth.cfg2(2)
    w = synth.make_weights(**kw)
    om = oracle_model(kw, w)
    mel = np.clip(np.random.RandomState(7).randn(2, 9, 80) * 1.5, -4, 4).astype(np.float32)
    got = om.upsample(mel)
    x = torch.from_numpy(mel)[:, None]                      # N,1,T,C
    for i, f in enumerate(kw['upsample_factor']):
        k = torch.from_numpy(w['wavenet/upsample%d/kernel' % i]).reshape(1, 1, f, 2)
        x = torch.nn.functional.conv_transpose2d(x, k, stride=(f, 1))[..., :80]
    assert got.shape == (2, 9 * 300, 80)
    np.testing.assert_allclose(got, x[:, 0].numpy(), atol=2e-6)
    np.testing.assert_allclose(got, np_oracle.NumpyWaveNet(**kw).__class__.create_upsample(_np(kw, w), mel), atol=2e-6)


def _np(kw, w):
    nn = np_oracle.NumpyWaveNet(**kw)
    nn.set_weights(w)
    return nn


def test_choice_equals_searchsorted_definition():
    # np.random.choice(arange(Q), p) == searchsorted(cumsum(float64 p)/sum, u, 'right')
    rs1, rs2 = np.random.RandomState(5), np.random.RandomState(5)
    gen = np.random.RandomState(11)
    for _ in range(500):
        p = gen.dirichlet(np.ones(256) * gen.choice([0.05, 1.0, 10.0])).astype(np.float32)
        p = (p / p.sum(dtype=np.float64)).astype(np.float32)
        a = rs1.choice(np.arange(256), p=p.astype(np.float64) / p.astype(np.float64).sum())
        u = rs2.random_sample()
        cdf = np.cumsum(p.astype(np.float64)); cdf /= cdf[-1]
        assert a == int(np.searchsorted(cdf, u, side='right'))


def test_temperature_one_leaves_prediction_unchanged():
    # the reference's runtime self-check, generate.py:227-228
    kw = synth.tiny_mulaw()
    w = synth.make_weights(**kw)
    om = oracle_model(kw, w)
    inp = make_inputs(kw, 6)
    _, lg = om.generate(6, inp['forced_full'][:, :6], inp['uniforms'], want_logits=True)
    for row in lg.reshape(-1, 256):
        pred = oracle.softmax_probs(row)
        np.testing.assert_allclose(pred, np_oracle.softmax_f64_to_f32(row[None])[0], rtol=2e-6, atol=1e-12)
        _, q = np_oracle.categorical_draw(pred[None], 1.0, [0.5])
        np.testing.assert_allclose(pred, q[0], atol=1e-5)
        assert abs(float(pred.astype(np.float64).sum()) - 1.0) < 1e-5


@pytest.mark.parametrize('name', ['tiny_mol', 'tiny_mulaw', 'cfg1', 'cfg2_n2'])
def test_oracle_reproduces_golden(name):
    from tests.golden.make_golden import CASES
    g = np.load(os.path.join(GOLD, name + '.npz'))
    fac, fkw, T = CASES[name]
    kw = fac(**fkw)
    w = synth.make_weights(**kw)
    om = oracle_model(kw, w)
    inp = make_inputs(kw, T)
    lc = om.upsample(inp['mel']) if 'mel' in inp else None
    plan = oracle.OrcPlan(*[int(v) for v in g['plan']])
    for tag, pl in (('kernel', plan), ('natural', oracle.OrcPlan.natural())):
        s, lg = om.generate(T, inp['x0'], inp['uniforms'], lc_up=lc, gc_ids=inp['gc_ids'], plan=pl, want_logits=True)
        assert np.array_equal(s, g['samples_' + tag])
        assert np.array_equal(lg[:, -4:], g['logits_tail_' + tag])
    if lc is not None:
        assert np.array_equal(lc[:, :7], g['lc_head'])
    # natural and kernel evaluation orders agree to float rounding (teacher forced)
    assert np.max(np.abs(g['tf_logits_kernel'] - g['tf_logits_natural'])) < 2e-5


@pytest.mark.parametrize('fac', [synth.tiny_mol, synth.tiny_mulaw])
def test_c_oracle_matches_numpy_restatement(fac):
    kw = fac()
    w = synth.make_weights(**kw)
    om, nn = oracle_model(kw, w), _np(kw, w)
    T = 64
    inp = make_inputs(kw, T)
    lc = om.upsample(inp['mel']) if 'mel' in inp else None
    s1, l1 = om.generate(T, inp['forced_full'], inp['uniforms'], lc_up=lc, gc_ids=inp['gc_ids'], want_logits=True)
    s2, l2 = np_oracle.generate(nn, T, inp['forced_full'], inp['uniforms'], lc_up=lc, gc_ids=inp['gc_ids'], want_logits=True)
    assert np.max(np.abs(l1 - l2)) < 1e-5
    if kw['scalar_input']:
        assert np.max(np.abs(s1 - s2)) < 1e-5
    else:
        assert np.mean(s1 != s2) <= 0.02       # identical up to cdf-boundary ties
    # free running, a few steps: the feedback path (sample -> causal queue) agrees
    s1 = om.generate(12, inp['x0'], inp['uniforms'][:, :12], lc_up=lc, gc_ids=inp['gc_ids'])
    s2 = np_oracle.generate(nn, 12, inp['x0'], inp['uniforms'][:, :12], lc_up=lc, gc_ids=inp['gc_ids'])
    assert np.max(np.abs(s1 - s2)) < 1e-4


def test_plan_changes_only_rounding():
    kw = synth.tiny_mol()
    w = synth.make_weights(**kw)
    om = oracle_model(kw, w)
    T = 40
    inp = make_inputs(kw, T)
    lc = om.upsample(inp['mel'])
    ref = om.generate(T, inp['forced_full'], inp['uniforms'], lc_up=lc, gc_ids=inp['gc_ids'], want_logits=True)[1]
    for plan in [oracle.OrcPlan(2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2), oracle.OrcPlan(4, 4, 4, 4, 4, 4, 4, 4, 8, 8, 4),
                 oracle.OrcPlan(1, 1, 16, 16, 4, 8, 16, 16, 32, 32, 8)]:
        got = om.generate(T, inp['forced_full'], inp['uniforms'], lc_up=lc, gc_ids=inp['gc_ids'], plan=plan, want_logits=True)[1]
        assert 0 < np.max(np.abs(got - ref)) < 2e-5


def test_priming_equals_forced_prefix():
    # generate.py:177-180: feeding seed[-rf:-1] with zero LC, then generating, is n_forced = len(seed) with
    # lc_shift = len(seed) - 1.  Rows are independent and the oracle is deterministic.
    kw = synth.tiny_mol()
    w = synth.make_weights(**kw)
    om = oracle_model(kw, w)
    rf = om.receptive_field
    T = rf + 30
    inp = make_inputs(kw, T)
    lc = om.upsample(inp['mel'])
    seed = inp['forced_full'][:, :rf]
    a = om.generate(T, seed, inp['uniforms'], lc_up=lc, lc_shift=rf - 1, gc_ids=inp['gc_ids'])
    b = om.generate(T, seed, inp['uniforms'], lc_up=lc, lc_shift=rf - 1, gc_ids=inp['gc_ids'])
    assert np.array_equal(a, b)
    # a different LC must not influence the priming steps' outputs (zero LC there)
    c = om.generate(rf - 1, seed[:, :rf - 1], inp['uniforms'][:, :rf - 1], lc_up=lc * 0 + 1, lc_shift=rf - 1, gc_ids=inp['gc_ids'])
    assert np.array_equal(a[:, :rf - 1], c)


def test_oracle_errors():
    kw = synth.tiny_mol()
    om = oracle.OracleModel(**kw)
    with pytest.raises(RuntimeError):
        om.set_weights({'wavenet/nonsense': np.zeros(3, np.float32)})
    with pytest.raises(RuntimeError):
        om.generate(4, np.zeros((2, 1), np.float32), np.full((2, 4, 11), 0.5, np.float32))   # no weights


@pytest.mark.parametrize('threads', [1, 3, 4])
def test_best_effort_cpu_organisation_is_bit_identical_to_the_port(threads):
    """oracle/wn_cpu_best.h (BASELINE.md B-cpu: weight-stationary threads, all rows together) against orc_generate: same samples and
    logits bit for bit, for the natural plan and a chunked one, MoL and mu-law heads, with and without conditioning, for
    thread counts that do and do not divide the channel counts."""
    from tacotron_wavenet_vocoder_korean_b200 import synth
    from tests.helpers import make_inputs, oracle_model
    mulaw_lc = dict(synth.tiny_mulaw(), local_condition_channels=20, upsample_factor=[2, 3], global_condition_channels=8,
                    global_condition_cardinality=3)
    chunked = oracle.OrcPlan.natural()
    chunked.M, chunked.Mt, chunked.t_cur, chunked.t_old, chunked.t_dense, chunked.t_skip, chunked.t_post1 = 2, 2, 4, 2, 2, 4, 4
    from tacotron_wavenet_vocoder_korean_b200.wavenet.model import plan_config
    kernel_plan = plan_from_dict(plan_config(148, **synth.cfg2(5))[0])          # the evaluation plan the B200 kernels report for cfg-2
    for kw, T, plan in ((synth.tiny_mol(), 120, None), (synth.tiny_mol(), 60, chunked), (synth.tiny_mulaw(), 80, None), (mulaw_lc, 60, None),
                        (synth.cfg2(5), 12, None), (synth.cfg2(5), 8, kernel_plan)):
        om = oracle_model(kw, synth.make_weights(**kw))
        inp = make_inputs(kw, T)
        lc = om.upsample(inp['mel']) if 'mel' in inp else None
        forced = inp['forced_full'][:, :7]                      # a primed start, then free running
        a = om.generate(T, forced, inp['uniforms'], lc_up=lc, gc_ids=inp['gc_ids'], want_logits=True, plan=plan)
        b = om.generate(T, forced, inp['uniforms'], lc_up=lc, gc_ids=inp['gc_ids'], want_logits=True, plan=plan, best_effort_threads=threads)
        assert np.array_equal(a[1], b[1]) and np.array_equal(a[0], b[0])


def test_best_effort_vector_gate_is_the_pinned_scalar_gate():
    """The 8-lane tanh32 * sigmoid32 of oracle/wn_cpu_best.h against the scalar pinned functions: identical bits on the range
    edges of exp32 (-87, 88, +-44 for tanh), zeros of both signs, denormals, infinities, and two million random inputs."""
    edge = np.array([0.0, -0.0, 1e-45, -1e-45, 1e-38, 43.999, 44.0, 44.00001, -44.0, -44.00001, 87.0, -87.0, 87.00001, -87.00001,
                     88.0, 88.5, -88.5, 100.0, -100.0, np.inf, -np.inf, 0.5, -0.5, 1.0, -1.0, 20.0, -20.0, 3e-8, -3e-8], np.float32)
    f, g = np.meshgrid(edge, edge)
    rs = np.random.RandomState(0)
    f = np.concatenate([f.ravel(), rs.randn(1000003).astype(np.float32) * 4, rs.uniform(-100, 100, 1000000).astype(np.float32)])
    g = np.concatenate([g.ravel(), rs.randn(1000003).astype(np.float32) * 4, rs.uniform(-100, 100, 1000000).astype(np.float32)])
    zv, zs = oracle.gate_probe(f, g)
    assert np.array_equal(zv.view(np.uint32), zs.view(np.uint32))


def test_best_effort_cpu_sixteen_threads_eight_column_slices():
    """16 threads on cfg-2: every thread owns 8 gated channels (one AVX2 vector), the 8-rows-at-a-time kernel plus a remainder row."""
    from tacotron_wavenet_vocoder_korean_b200 import synth
    from tests.helpers import make_inputs, oracle_model
    kw = synth.cfg2(9)
    om = oracle_model(kw, synth.make_weights(**kw))
    T = 6
    inp = make_inputs(kw, T)
    lc = om.upsample(inp['mel'])
    a = om.generate(T, inp['x0'], inp['uniforms'], lc_up=lc, gc_ids=inp['gc_ids'], want_logits=True)
    b = om.generate(T, inp['x0'], inp['uniforms'], lc_up=lc, gc_ids=inp['gc_ids'], want_logits=True, best_effort_threads=16)
    assert np.array_equal(a[1], b[1]) and np.array_equal(a[0], b[0])
