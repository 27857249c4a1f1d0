"""GPU parity tests: the CUDA path (through the C ABI / WaveNetModel) against the CPU oracle on the
same seeded inputs, against the committed golden fixtures, and -- at BASELINE.json's full sizes --
through size-independent properties.  Tolerances: every output is compared BIT-EXACT (the oracle is run
with the evaluation plan the kernel reports); the 1e-4 bound of north_star is additionally checked against
the oracle's natural evaluation order."""
import ctypes as C
import os

import numpy as np
import pytest
import torch

import oracle
from tests.helpers import make_inputs, oracle_model, plan_from_dict
from tacotron_wavenet_vocoder_korean_b200 import _lib, synth
from tacotron_wavenet_vocoder_korean_b200.wavenet import WaveNetModel, mu_law_encode, mu_law_decode

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def build(kw, **extra):
    w = synth.make_weights(**kw)
    net = WaveNetModel(train_mode=False, **kw, **extra)
    net.load_state_dict(w)
    return net, w


def run_both(kw, T, teacher=False, temperature=1.0, lc_shift=0, forced=None, **extra):
    net, w = build(kw, **extra)
    om = oracle_model(kw, w)
    inp = make_inputs(kw, T)
    lc_o = lc_t = None
    if 'mel' in inp:
        lc_t = net.create_upsample(inp['mel'])
        lc_o = om.upsample(inp['mel'])
        assert np.array_equal(lc_o, lc_t.cpu().numpy()), 'create_upsample differs'
    if forced is None:
        forced = inp['forced_full'] if teacher else inp['x0']
    s, lg = net.generate(T, forced, inp['uniforms'], lc_up=lc_t, lc_shift=lc_shift, gc_ids=inp['gc_ids'],
                         temperature=temperature, want_logits=True)
    plan = plan_from_dict(net.plan())
    so, lo = om.generate(T, forced, inp['uniforms'], lc_up=lc_o, lc_shift=lc_shift, gc_ids=inp['gc_ids'],
                         temperature=temperature, plan=plan, want_logits=True)
    return net, om, inp, (s.cpu().numpy(), lg.cpu().numpy()), (so, lo), (lc_t, lc_o)


def assert_exact(got, exp):
    (s, lg), (so, lo) = got, exp
    if not np.array_equal(lo, lg):
        bad = np.argwhere(lo != lg)[0]
        raise AssertionError('logits differ first at (row, step, channel) %s: oracle %r kernel %r (max abs diff %g)'
                             % (bad, lo[tuple(bad)], lg[tuple(bad)], np.abs(lo - lg).max()))
    if not np.array_equal(so, s):
        bad = np.argwhere(so != s)[0]
        raise AssertionError('samples differ first at (row, step) %s: oracle %r kernel %r' % (bad, so[tuple(bad)], s[tuple(bad)]))


@pytest.mark.parametrize('fac,T', [(synth.tiny_mol, 150), (synth.tiny_mulaw, 300)])
@pytest.mark.parametrize('teacher', [False, True])
def test_tiny_bit_exact(fac, T, teacher):
    _, _, _, got, exp, _ = run_both(fac(), T, teacher=teacher)
    assert_exact(got, exp)


@pytest.mark.parametrize('fac', [synth.tiny_mol, synth.tiny_mulaw])
@pytest.mark.parametrize('M,Mt', [(1, 1), (2, 2), (4, 4), (2, 8), (4, 1)])
def test_every_split_bit_exact(fac, M, Mt):
    # every layer / tail split is a different evaluation plan and a different CTA topology
    kw = fac()
    if kw['skip_channels'] // Mt < 4:
        pytest.skip('slice too small')
    net, _, _, got, exp, _ = run_both(kw, 120, force_M=M, force_Mt=Mt)
    assert net.plan()['M'] == M and net.plan()['Mt'] == Mt
    assert_exact(got, exp)


def test_cfg1_full_length_integer_samples_bit_exact():
    # BASELINE configs[0]: 10-layer mu-law 256, 0.5 s @ 16 kHz = 8000 steps, unconditioned, free running
    kw = synth.cfg1()
    net, om, inp, got, exp, _ = run_both(kw, 8000)
    assert_exact(got, exp)
    s = got[0]
    assert s.min() >= 0 and s.max() <= 255 and np.all(s == np.round(s))
    assert len(np.unique(s)) > 100                       # a real draw, not a stuck value
    # north_star tolerance vs the natural evaluation order (teacher forcing with the kernel's own samples)
    forced = np.concatenate([inp['x0'], s[:, :-1]], axis=1)
    _, lnat = om.generate(8000, forced, inp['uniforms'], want_logits=True)
    assert np.max(np.abs(lnat - got[1])) < 1e-4


def test_cfg2_batch8_bit_exact_and_tolerance():
    # BASELINE configs[1]: 30 layers, R=D=128, S=512, MoL-10, mel + speaker conditioned, batch 8
    kw = synth.cfg2(8)
    net, om, inp, got, exp, (lc_t, lc_o) = run_both(kw, 260)
    info = net.info()
    assert (info['grid'], info['M'], info['Mt']) == (137, 4, 16) and info['weights_in_global'] == 0 and info['static_shape'] == 1
    assert_exact(got, exp)
    assert np.all(np.abs(got[0]) <= 1.0)
    forced = np.concatenate([inp['x0'], got[0][:, :-1]], axis=1)
    _, lnat = om.generate(260, forced, inp['uniforms'], lc_up=lc_o, gc_ids=inp['gc_ids'], want_logits=True)
    assert np.max(np.abs(lnat - got[1])) < 1e-4          # float MoL logits within 1e-4 (north_star)


def _oracle_threads(n):
    oracle.lib().orc_set_threads(int(n))


@pytest.mark.parametrize('rows,T', [(2, 1200), (12, 1100)])
def test_cfg2_dilation_512_rings_free_running_bit_exact(rows, T):
    """dilation 512 first reads a non-zero delayed tap at t = 512.  Free-running cfg-2 (the shape the bench
    runs: cluster path, 2 rows; 12 rows = the many-row regime) against the oracle well past that point, bit for bit."""
    kw = synth.cfg2(rows)
    _oracle_threads(8)
    try:
        net, _, _, got, exp, _ = run_both(kw, T)
    finally:
        _oracle_threads(1)
    assert net.info()['static_shape'] == 1
    assert_exact(got, exp)


def test_cfg2_teacher_forced_two_receptive_fields_bit_exact():
    """>= 2 * receptive field (3101) teacher-forced steps at the cfg-2 shape: every ring has wrapped at least six times
    (SURVEY 8c asks for >= 2 rf); logits bit-exact vs the oracle with the kernel's plan, <= 1e-4 vs its natural order."""
    kw = synth.cfg2(1)
    T = 2 * 3101 + 98
    net, om, inp, got, exp, (lc_t, lc_o) = run_both(kw, T, teacher=True)
    assert net.receptive_field == 3101
    assert_exact(got, exp)
    _, lnat = om.generate(T, inp['forced_full'], inp['uniforms'], lc_up=lc_o, gc_ids=inp['gc_ids'], want_logits=True)
    assert np.max(np.abs(lnat - got[1])) < 1e-4


def test_cluster_path_equals_single_kernel_path():
    """The round-2 cluster / DSMEM path (wn_kernel_v2.cuh: 15 clusters of 8 CTAs + tail kernel) and the round-1 single
    cooperative kernel implement the same evaluation plan: identical bits, for full, partial and ragged batches."""
    kw = synth.cfg2(8)
    a, _ = build(kw)
    b, _ = build(kw, cluster=False)
    assert a.info()['cluster_path'] >= 1 and b.info()['cluster_path'] == 0 and a.plan() == b.plan()
    T = 700
    inp = make_inputs(kw, T)
    lc = a.create_upsample(inp['mel'])
    for rows, T_row in ((8, None), (3, None), (8, [700, 17, 0, 255, 700, 1, 699, 64])):
        args = (T, inp['x0'][:rows], inp['uniforms'][:rows])
        kws = dict(lc_up=lc[:rows], gc_ids=inp['gc_ids'][:rows], want_logits=True, T_row=T_row)
        sa, la = a.generate(*args, **kws)
        sb, lb = b.generate(*args, **kws)
        for r in range(rows):
            n = T if T_row is None else T_row[r]
            assert torch.equal(sa[r, :n], sb[r, :n]) and torch.equal(la[r, :n], lb[r, :n])


def test_folded_upsample_equals_materialised_condition():
    """north_star: "mel local-conditioning upsample staged via TMA".  Passing mel frames (wn_generate_args.mel_dev) makes the
    layer CTAs evaluate create_upsample (model.py:102-111) per step from TMA-staged frames; samples and logits are
    bit-identical to feeding the materialised (rows, T, 80) tensor, also with priming (lc_shift) and ragged rows, and
    bit-identical to the oracle (which upsamples on the host)."""
    kw = synth.cfg2(4)
    net, w = build(kw)
    T = 1500                                                    # five mel frames
    inp = make_inputs(kw, T)
    lc = net.create_upsample(inp['mel'])
    a = net.generate(T, inp['x0'], inp['uniforms'], lc_up=lc, gc_ids=inp['gc_ids'], want_logits=True)
    b = net.generate(T, inp['x0'], inp['uniforms'], mel=inp['mel'], gc_ids=inp['gc_ids'], want_logits=True)
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])
    T_row = [1500, 299, 301, 0]
    a = net.generate(T, inp['forced_full'][:, :350], inp['uniforms'], lc_up=lc, lc_shift=349, gc_ids=inp['gc_ids'], T_row=T_row)
    b = net.generate(T, inp['forced_full'][:, :350], inp['uniforms'], mel=inp['mel'], lc_shift=349, gc_ids=inp['gc_ids'], T_row=T_row)
    for r, n in enumerate(T_row):
        assert torch.equal(a[r, :n], b[r, :n])
    om = oracle_model(kw, w)
    _oracle_threads(4)
    try:
        so = om.generate(600, inp['x0'], inp['uniforms'][:, :600], lc_up=om.upsample(inp['mel'])[:, :600], gc_ids=inp['gc_ids'],
                         plan=plan_from_dict(net.plan()))
    finally:
        _oracle_threads(1)
    s = net.generate(600, inp['x0'], inp['uniforms'][:, :600], mel=inp['mel'], gc_ids=inp['gc_ids']).cpu().numpy()
    assert np.array_equal(s, so)
    # models without the 3-stage upsampler of cfg-2 / other kernels materialise internally: same results
    kw = synth.tiny_mol()
    net, _ = build(kw)
    inp = make_inputs(kw, 120)
    a = net.generate(120, inp['x0'], inp['uniforms'], lc_up=net.create_upsample(inp['mel']), gc_ids=inp['gc_ids'])
    b = net.generate(120, inp['x0'], inp['uniforms'], mel=inp['mel'], gc_ids=inp['gc_ids'])
    assert torch.equal(a, b)


def test_fast_activation_within_north_star_tolerance():
    """WN_FLAG_FAST_ACT (ex2.approx / rcp.approx gate, scalar-input path only): teacher-forced logits within 1e-4 of the
    oracle (north_star's bound for float MoL logits) over more than one receptive field, and of the reference's own run."""
    kw = synth.cfg2(2)
    net, w = build(kw, fast_act=True)
    assert net.info()['fast_act'] == 1
    om = oracle_model(kw, w)
    T = 3300
    inp = make_inputs(kw, T)
    lc = net.create_upsample(inp['mel'])
    _oracle_threads(2)
    try:
        so, lo = om.generate(T, inp['forced_full'], inp['uniforms'], lc_up=lc.cpu().numpy(), gc_ids=inp['gc_ids'], want_logits=True)
    finally:
        _oracle_threads(1)
    s, lg = net.generate(T, inp['forced_full'], inp['uniforms'], lc_up=lc, gc_ids=inp['gc_ids'], want_logits=True)
    assert np.abs(lg.cpu().numpy() - lo).max() < 1e-4 and np.abs(s.cpu().numpy() - so).max() < 1e-4
    g = np.load(os.path.join(GOLD, 'ref_cfg2.npz'))                      # the reference's own wavenet/model.py, 640 steps
    Tr = g['outputs'].shape[1]
    inp = make_inputs(kw, Tr)
    lc = net.create_upsample(inp['mel'])
    s, lg = net.generate(Tr, inp['forced_full'], inp['uniforms'], lc_up=lc, gc_ids=inp['gc_ids'], want_logits=True)
    assert np.abs(lg.cpu().numpy() - g['raw_output']).max() < 1e-4 and np.abs(s.cpu().numpy() - g['outputs'][:, :, 0]).max() < 1e-4
    # free running: a valid waveform (the draw is chaotic in the last bits, so no sample-level comparison)
    f = net.generate(2000, inp['x0'], make_inputs(kw, 2000)['uniforms'], lc_up=net.create_upsample(make_inputs(kw, 2000)['mel']), gc_ids=inp['gc_ids']).cpu().numpy()
    assert np.all(np.isfinite(f)) and np.all(np.abs(f) <= 1.0) and f.std() > 1e-3


@pytest.mark.parametrize('fac,T,shape', [(lambda: synth.cfg2(3), 150, 1), (synth.cfg1, 600, 2), (lambda: synth.cfg_hparams_default(2), 200, 3)])
def test_runtime_shaped_kernel_equals_specialised_kernel(fac, T, shape):
    # the compile-time specialised instantiations and the generic kernel implement the same plan
    kw = fac()
    net_s, w = build(kw)
    net_g, _ = build(kw, generic_kernel=True)
    assert net_s.info()['static_shape'] == shape and net_g.info()['static_shape'] == 0
    assert net_s.plan() == net_g.plan()
    inp = make_inputs(kw, T)
    lc = net_s.create_upsample(inp['mel']) if 'mel' in inp else None
    a = net_s.generate(T, inp['x0'], inp['uniforms'], lc_up=lc, gc_ids=inp['gc_ids'], want_logits=True)
    b = net_g.generate(T, inp['x0'], inp['uniforms'], lc_up=lc, gc_ids=inp['gc_ids'], want_logits=True)
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])
    om = oracle_model(kw, w)
    so = om.generate(T, inp['x0'], inp['uniforms'], lc_up=lc.cpu().numpy() if lc is not None else None,
                     gc_ids=inp['gc_ids'], plan=plan_from_dict(net_g.plan()))
    assert np.array_equal(so, b[0].cpu().numpy())


def test_dual_homed_mailboxes_do_not_change_results():
    # die-aware (dual-homed) mailboxes are a placement optimisation only
    kw = synth.cfg2(2)
    net_a, _ = build(kw)
    net_b, _ = build(kw, die_aware=False)
    assert net_b.info()['die_aware'] == 0
    inp = make_inputs(kw, 200)
    lc = net_a.create_upsample(inp['mel'])
    a = net_a.generate(200, inp['x0'], inp['uniforms'], lc_up=lc, gc_ids=inp['gc_ids'])
    b = net_b.generate(200, inp['x0'], inp['uniforms'], lc_up=lc, gc_ids=inp['gc_ids'])
    assert torch.equal(a, b)


@pytest.mark.parametrize('cluster', [True, False])
def test_many_rows_variant_bit_exact(cluster):
    # many rows in flight (round-1 path: >= 10 rows dispatch to the warp-specialised layer CTA, wn_kernel_ws.cuh); same bits
    kw = synth.cfg2(12)
    net, w = build(kw, cluster=cluster)
    T = 150
    inp = make_inputs(kw, T)
    lc = net.create_upsample(inp['mel'])
    a = net.generate(T, inp['x0'], inp['uniforms'], lc_up=lc, gc_ids=inp['gc_ids'], want_logits=True)         # 12 rows: ws
    b = net.generate(T, inp['x0'][:5], inp['uniforms'][:5], lc_up=lc[:5], gc_ids=inp['gc_ids'][:5], want_logits=True)   # 5 rows: single group
    assert torch.equal(a[0][:5], b[0]) and torch.equal(a[1][:5], b[1])
    om = oracle_model(kw, w)
    so, lo = om.generate(T, inp['x0'], inp['uniforms'], lc_up=lc.cpu().numpy(), gc_ids=inp['gc_ids'],
                         plan=plan_from_dict(net.plan()), want_logits=True)
    assert_exact((a[0].cpu().numpy(), a[1].cpu().numpy()), (so, lo))
    # ragged rows and priming through the same variant
    T_row = [150, 3, 0, 77, 150, 150, 9, 150, 150, 1, 150, 60]
    s = net.generate(T, inp['x0'], inp['uniforms'], lc_up=lc, gc_ids=inp['gc_ids'], T_row=T_row).cpu().numpy()
    for r, tr in enumerate(T_row):
        assert np.array_equal(s[r, :tr], so[r, :tr])


def test_shared_ring_build_equals_private_rings():
    """From 16 rows on the cluster path keeps ONE tagged dilation ring per layer (wn_kernel_v2.cuh, SR) instead of one private copy
    per sibling CTA.  Same bits as the private-ring build on the same 16-row job past the d = 512 wrap, in both directions on
    one handle (the shared ring is zeroed per launch, the private rings are written before they are read), and when forced at 8 rows."""
    kw = synth.cfg2(16)
    net, w = build(kw)
    T = 700
    inp = make_inputs(kw, T)
    lc = net.create_upsample(inp['mel'])

    def run(rows, flag):
        if flag is None:
            os.environ.pop('WN_SHARED_RING', None)
        else:
            os.environ['WN_SHARED_RING'] = flag
        try:
            return net.generate(T, inp['x0'][:rows], inp['uniforms'][:rows], lc_up=lc[:rows], gc_ids=inp['gc_ids'][:rows], want_logits=True)
        finally:
            os.environ.pop('WN_SHARED_RING', None)
    shared = run(16, None)                  # default at 16 rows: shared ring
    private = run(16, '0')
    again = run(16, '1')
    assert torch.equal(shared[0], private[0]) and torch.equal(shared[1], private[1])
    assert torch.equal(shared[0], again[0]) and torch.equal(shared[1], again[1])
    eight = run(8, None)                    # default at 8 rows: private rings (the benchmark path)
    eight_sr = run(8, '1')
    assert torch.equal(eight[0], eight_sr[0]) and torch.equal(eight[1], eight_sr[1])
    assert torch.equal(eight[0], shared[0][:8])
    om = oracle_model(kw, w)
    _oracle_threads(8)
    try:
        so = om.generate(T, inp['x0'], inp['uniforms'], lc_up=lc.cpu().numpy(), gc_ids=inp['gc_ids'], plan=plan_from_dict(net.plan()))
    finally:
        _oracle_threads(1)
    assert np.array_equal(so, shared[0].cpu().numpy())
    # ragged rows (some finished early, one never started) through both builds
    T_row = [700, 3, 0, 77, 700, 513, 9, 700, 650, 1, 700, 60, 700, 512, 700, 300]
    outs = []
    for flag in ('1', '0'):
        os.environ['WN_SHARED_RING'] = flag
        try:
            outs.append(net.generate(T, inp['x0'], inp['uniforms'], lc_up=lc, gc_ids=inp['gc_ids'], T_row=T_row).cpu().numpy())
        finally:
            os.environ.pop('WN_SHARED_RING', None)
    for r, tr in enumerate(T_row):
        assert np.array_equal(outs[0][r, :tr], so[r, :tr]) and np.array_equal(outs[1][r, :tr], so[r, :tr])


def test_hparams_default_model_bit_exact():
    # the reference's own defaults (hparams.py:59-79): 50 layers, R=D=32, scalar input, lc + gc
    _, _, _, got, exp, _ = run_both(synth.cfg_hparams_default(2), 400)
    assert_exact(got, exp)


@pytest.mark.parametrize('name', ['tiny_mol', 'tiny_mulaw', 'cfg1', 'cfg2_n2'])
def test_matches_committed_golden(name):
    from tests.golden.make_golden import CASES
    g = np.load(os.path.join(GOLD, name + '.npz'))
    fac, fkw, T = CASES[name]
    kw = fac(**fkw)
    net, _ = build(kw)
    inp = make_inputs(kw, T)
    assert [net.plan()[k] for k, _ in oracle.OrcPlan._fields_] == [int(v) for v in g['plan']]
    lc = net.create_upsample(inp['mel']) if 'mel' in inp else None
    s, lg = net.generate(T, inp['x0'], inp['uniforms'], lc_up=lc, gc_ids=inp['gc_ids'], want_logits=True)
    assert np.array_equal(s.cpu().numpy(), g['samples_kernel'])
    assert np.array_equal(lg[:, -4:].cpu().numpy(), g['logits_tail_kernel'])
    _, lg = net.generate(48, inp['forced_full'][:, :48], inp['uniforms'][:, :48], lc_up=lc, gc_ids=inp['gc_ids'], want_logits=True)
    assert np.array_equal(lg.cpu().numpy(), g['tf_logits_kernel'])
    assert np.max(np.abs(lg.cpu().numpy() - g['tf_logits_natural'])) < 1e-4
    if lc is not None:
        assert np.array_equal(lc[:, :7].cpu().numpy(), g['lc_head'])


def test_temperature_changes_draw_and_stays_exact():
    _, _, _, got, exp, _ = run_both(synth.tiny_mulaw(), 200, temperature=0.7)
    assert_exact(got, exp)
    _, _, _, got1, _, _ = run_both(synth.tiny_mulaw(), 200, temperature=1.0)
    assert not np.array_equal(got[0], got1[0])


def test_priming_with_seed_path():
    # generate.py:170-180: rf-1 priming steps with zero LC, then generation from seed[-1]
    kw = synth.tiny_mol()
    rf = WaveNetModel.calculate_receptive_field(2, kw['dilations'], True, kw['initial_filter_width'])
    T = rf - 1 + 90
    inp = make_inputs(kw, T)
    seed = inp['forced_full'][:, :rf]
    _, _, _, got, exp, _ = run_both(kw, T, forced=seed, lc_shift=rf - 1)
    assert_exact(got, exp)


def test_ragged_rows_and_partial_batch():
    kw = synth.tiny_mol(batch_size=4)
    net, w = build(kw)
    om = oracle_model(kw, w)
    T = 90
    inp = make_inputs(kw, T)
    lc = net.create_upsample(inp['mel'])
    T_row = [90, 17, 0, 55]
    s = net.generate(T, inp['x0'], inp['uniforms'], lc_up=lc, gc_ids=inp['gc_ids'], T_row=T_row).cpu().numpy()
    so = om.generate(T, inp['x0'], inp['uniforms'], lc_up=lc.cpu().numpy(), gc_ids=inp['gc_ids'], plan=plan_from_dict(net.plan()))
    for b, tr in enumerate(T_row):
        assert np.array_equal(s[b, :tr], so[b, :tr])
    # fewer rows than batch_size: rows are independent, so row b equals the full-batch result
    s2 = net.generate(T, inp['x0'][:2], inp['uniforms'][:2], lc_up=lc[:2], gc_ids=inp['gc_ids'][:2]).cpu().numpy()
    assert np.array_equal(s2, so[:2])
    # empty job
    assert net.generate(0, inp['x0'], inp['uniforms'][:, :0], lc_up=lc, gc_ids=inp['gc_ids']).shape == (4, 0)


def test_host_buffer_entry_point_equals_device_entry_point():
    kw = synth.tiny_mol()
    net, _ = build(kw)
    T = 120
    inp = make_inputs(kw, T)
    lc = net.create_upsample(inp['mel'])
    dev = net.generate(T, inp['x0'], inp['uniforms'], lc_up=lc, gc_ids=inp['gc_ids']).cpu().numpy()
    host = net.generate_host(T, inp['x0'], inp['uniforms'], mel=inp['mel'], gc_ids=inp['gc_ids'])
    assert np.array_equal(dev, host)


def test_predict_proba_incremental_is_a_true_single_step():
    """the drop-in predict_proba_incremental must be an O(1) state-carrying step (wn_step, persistent device
    queues), usable by a caller that keeps the reference's own per-sample loop (generate.py:202-233).  Drive that loop
    shape for more than a receptive field and compare with the fused wn_generate: identical ids / samples."""
    import time
    # --- one-hot model: probabilities out, host-style categorical draw (generate.py:219-231) fed back --------------
    kw = synth.tiny_mulaw()
    net, w = build(kw)
    om = oracle_model(kw, w)
    T = net.receptive_field + 60
    inp = make_inputs(kw, T)
    fused, logits = net.generate(T, inp['x0'], inp['uniforms'], want_logits=True)
    net.reset_incremental()
    x = torch.as_tensor(inp['x0'], dtype=torch.float32).cuda()
    ids = []
    for t in range(T):
        proba, draw = net.predict_proba_incremental(x, uniforms=inp['uniforms'][:, t], return_draw=True)
        assert proba.shape == (2, 256)
        if t < 5 or t == T - 1:
            ref = np.stack([oracle.softmax_probs(r) for r in logits[:, t].cpu().numpy()])
            assert np.array_equal(proba.cpu().numpy(), ref)                         # float32(softmax(float64(.))), model.py:243
            np.testing.assert_allclose(proba.sum(dim=1).cpu().numpy(), 1.0, atol=1e-5)   # generate.py:227 invariant
        ids.append(draw)
        x = draw.reshape(2, 1)
    assert torch.equal(torch.stack(ids, dim=1), fused)
    # --- the benchmark model: scalar input, mel + speaker conditioned, draws inside the node (mixture.py:84-114) ------
    kw = synth.cfg2(2)
    net, w = build(kw)
    T = net.receptive_field + 120                                                  # 3221 calls
    inp = make_inputs(kw, T)
    lc = net.create_upsample(inp['mel'])
    fused = net.generate(T, inp['x0'], inp['uniforms'], mel=inp['mel'], gc_ids=inp['gc_ids'])
    uni = torch.as_tensor(inp['uniforms']).cuda()
    x = torch.as_tensor(inp['x0'], dtype=torch.float32).cuda()
    out, stamps = [], []
    for t in range(T):
        if t in (100, 200, T - 200, T - 100):
            torch.cuda.synchronize()
            stamps.append(time.perf_counter())
        x = net.predict_proba_incremental(x, lc[:, t], inp['gc_ids'], uniforms=uni[:, t])
        out.append(x[:, 0])
    assert torch.equal(torch.stack(out, dim=1), fused)
    early, late = stamps[1] - stamps[0], stamps[3] - stamps[2]
    assert late < 2.0 * early + 0.05, (early, late)                                 # O(1) per call: no history replay
    # queue_initializer starts over
    net.reset_incremental()
    x0 = torch.as_tensor(inp['x0'], dtype=torch.float32).cuda()
    again = net.predict_proba_incremental(x0, lc[:, 0], inp['gc_ids'], uniforms=uni[:, 0])
    assert torch.equal(again[:, 0], fused[:, 0])


def test_mu_law_codec_on_device():
    g = np.load(os.path.join(GOLD, 'codec.npz'))
    enc = mu_law_encode(torch.from_numpy(g['grid']).cuda(), 256).cpu().numpy()
    assert np.array_equal(enc, g['enc'])                      # integer codes: bit-exact (pinned log1p32 on both sides)
    assert np.array_equal(enc, oracle.mu_law_encode(g['grid'], 256))
    dec = mu_law_decode(torch.arange(256, dtype=torch.float32).cuda(), 256, True).cpu().numpy()
    np.testing.assert_allclose(dec, g['dec_q'], atol=1e-4, rtol=1e-5)
    assert np.array_equal(mu_law_encode(torch.from_numpy(dec).cuda(), 256).cpu().numpy(), np.arange(256))
    dec_c = mu_law_decode(torch.linspace(-1, 1, 513).cuda(), 256, False).cpu().numpy()
    np.testing.assert_allclose(dec_c, g['dec_c'], atol=1e-4, rtol=1e-5)


def test_mixture_module_sample_and_loss_on_device():
    """wavenet/mixture.py as tensor-level entry points (wn_mol_sample / wn_mol_loss): the draw is bit-identical to the oracle's
    mol_draw and within 2e-5 of the reference's own draws on the reference's own logits (ref_mol.npz); the loss follows the fp64
    numpy evaluation of mixture.py:27-81 in every tf.where branch at the fp32 tolerance of the oracle's own fp32 evaluation."""
    from oracle import train_oracle as to
    from tacotron_wavenet_vocoder_korean_b200.wavenet.mixture import sample_from_discretized_mix_logistic, discretized_mix_logistic_loss
    g = np.load(os.path.join(GOLD, 'ref_mol.npz'))
    kw = synth.tiny_mol()
    inp = make_inputs(kw, g['outputs'].shape[1])
    s = sample_from_discretized_mix_logistic(torch.from_numpy(g['raw_output']).cuda(), uniforms=torch.from_numpy(inp['uniforms']).cuda()).cpu().numpy()
    assert np.array_equal(s, oracle.mol_sample(g['raw_output'], inp['uniforms']))
    assert np.abs(s - g['outputs'][:, :, 0]).max() < 2e-5
    rs = np.random.RandomState(7)
    B, T, K = 3, 1000, 10
    y = rs.randn(B, T, 3 * K).astype(np.float32)
    y[..., 2 * K:] = rs.uniform(-9, 1, (B, T, K))
    y[0, :50, 2 * K:] = -40.0                                  # below log_scale_min: tf.maximum clamp
    u = rs.uniform(1e-5, 1 - 1e-5, (B, T, K + 1)).astype(np.float32)
    u[1, :10, :K], u[2, :10, K] = 1e-5, 1 - 1e-5                # extremes of the uniform range
    s = sample_from_discretized_mix_logistic(torch.from_numpy(y).cuda(), uniforms=torch.from_numpy(u).cuda()).cpu().numpy()
    assert np.array_equal(s, oracle.mol_sample(y, u))
    assert s.min() >= -1.0 and s.max() <= 1.0 and (s == 1.0).any() and (s == -1.0).any()      # the clip is exercised
    r = sample_from_discretized_mix_logistic(torch.from_numpy(y).cuda())                          # own uniforms: in range, not constant
    assert r.shape == (B, T) and float(r.min()) >= -1.0 and float(r.max()) <= 1.0 and float(r.std()) > 0.05
    # loss: the case of tests/test_train_oracle.py::test_mol_loss_matches_float64_numpy_in_every_branch
    rs = np.random.RandomState(0)
    B, T = 2, 400
    y_hat = rs.randn(B, T, 3 * K).astype(np.float32)
    y_hat[..., 2 * K:] = rs.uniform(-9, -1, (B, T, K))
    y_hat[0, :40, 2 * K:] = -40.0
    tgt = rs.uniform(-1, 1, (B, T, 1)).astype(np.float32)
    tgt[0, :20], tgt[1, :20] = -1.0, 1.0
    y_hat[1, 100:140, K:2 * K] = 30.0                            # cdf_delta <= 1e-5: log-pdf branch
    for nc in (256, 65536):
        ref = to.mol_loss_np(y_hat, tgt, num_class=nc)
        got = discretized_mix_logistic_loss(torch.from_numpy(y_hat).cuda(), torch.from_numpy(tgt).cuda(), num_class=nc, reduce=False).cpu().numpy()
        np.testing.assert_allclose(got, ref, rtol=5e-3, atol=1e-3)   # fp32: cdf_plus - cdf_min cancels (same bound as the fp32 oracle)
        ref32 = to.mol_loss(torch.from_numpy(y_hat), torch.from_numpy(tgt), num_class=nc).numpy()
        assert np.abs(got - ref).max() <= 3 * np.abs(ref32 - ref).max() + 1e-5          # not noisier than the fp32 oracle itself
        tot = float(discretized_mix_logistic_loss(torch.from_numpy(y_hat).cuda(), torch.from_numpy(tgt).cuda(), num_class=nc))
        assert abs(tot - float(got.astype(np.float64).sum())) < 1e-6 * abs(tot)
    with pytest.raises(AssertionError):
        sample_from_discretized_mix_logistic(torch.zeros(1, 4, 31).cuda())


def test_argument_errors_are_loud():
    kw = synth.tiny_mol()
    net, _ = build(kw)
    inp = make_inputs(kw, 8)
    lc = net.create_upsample(inp['mel'])
    with pytest.raises(RuntimeError, match='gc_ids'):
        net.generate(8, inp['x0'], inp['uniforms'], lc_up=lc)                        # generate.py:72-77
    with pytest.raises(ValueError):
        net.generate(8, inp['x0'], inp['uniforms'][:, :, :5], lc_up=lc, gc_ids=inp['gc_ids'])
    with pytest.raises(RuntimeError):
        net.generate(8, inp['x0'], inp['uniforms'], lc_up=lc, gc_ids=[0, 99])
    fresh = WaveNetModel(train_mode=False, **kw)
    with pytest.raises(RuntimeError):
        fresh.generate(8, inp['x0'], inp['uniforms'], lc_up=lc, gc_ids=inp['gc_ids'])   # not loaded
    with pytest.raises(RuntimeError, match='missing weight'):
        fresh.load_state_dict({k: v for k, v in synth.make_weights(**kw).items() if 'skip/kernel' not in k})


def test_generate_cli_end_to_end(tmp_path):
    """generate.py entry point (reference CLI, generate.py:38-79): params.json + mel .npy in, wav files out."""
    import json
    from scipy.io import wavfile
    from tacotron_wavenet_vocoder_korean_b200 import generate as gen
    from tacotron_wavenet_vocoder_korean_b200.hparams import hparams
    saved = hparams.values()
    try:
        ck = tmp_path / 'ckpt'
        ck.mkdir()
        kw = synth.cfg_hparams_default(2)
        json.dump({'dilations': kw['dilations'][:12], 'residual_channels': 32, 'skip_channels': 64, 'sample_rate': 24000},
                  open(ck / 'params.json', 'w'))
        mel = np.clip(np.random.RandomState(4).randn(3, 80) * 1.5, -4, 4).astype(np.float32)
        np.save(tmp_path / 'mel.npy', mel)
        argv = [str(ck), '--mel', str(tmp_path / 'mel.npy'), '--gc_cardinality', '2', '--gc_id', '1', '--batch_size', '2',
                '--logdir', str(tmp_path / 'log'), '--seed', '7', '--synthetic_weights']
        wav = gen.main(argv)
        assert wav.shape == (2, 3 * 300) and np.all(np.abs(wav) <= 1.0)
        assert not np.array_equal(wav[0], wav[1])                   # rows share mel/gc but draw different uniforms
        files = sorted((tmp_path / 'log' / 'generate').glob('*/test-*.wav'))
        assert len(files) == 2
        sr, data = wavfile.read(files[0])
        assert sr == 24000 and data.dtype == np.int16 and len(data) == 900 and np.abs(data).max() == 32767
        # same seed -> same audio (the reference is unseeded; --seed is the reproducibility hook)
        wav2 = gen.main(argv)
        assert np.array_equal(wav, wav2)
        with pytest.raises(ValueError):
            gen.main([str(ck), '--mel', str(tmp_path / 'mel.npy'), '--synthetic_weights'])      # generate.py:72-77
    finally:
        for k, v in saved.items():
            setattr(hparams, k, v)


def test_full_size_properties_cfg2():
    """BASELINE configs[1] at full length (8 utterances x 2 s @ 24 kHz = 48000 steps), checked through
    properties that do not need a 48000-step CPU run: determinism, causality (prefix property), batch-row
    independence, and an oracle check on a bounded prefix."""
    kw = synth.cfg2(8)
    net, w = build(kw)
    T = 48000
    inp = make_inputs(kw, T, t_mel=160)
    assert inp['mel'].shape == (8, 160, 80)
    lc = net.create_upsample(inp['mel'])
    assert lc.shape == (8, 48000, 80)
    a = net.generate(T, inp['x0'], inp['uniforms'], lc_up=lc, gc_ids=inp['gc_ids'])
    b = net.generate(T, inp['x0'], inp['uniforms'], lc_up=lc, gc_ids=inp['gc_ids'])
    assert torch.equal(a, b)                                                   # deterministic
    a = a.cpu().numpy()
    assert np.all(np.isfinite(a)) and np.all(np.abs(a) <= 1.0) and a.std() > 1e-3
    k = 3000
    pre = net.generate(k, inp['x0'], inp['uniforms'][:, :k], lc_up=lc, gc_ids=inp['gc_ids']).cpu().numpy()
    assert np.array_equal(pre, a[:, :k])                                        # causal: prefix of the long run
    # rows never interact (model.py:112-167 is row-wise): row 5 alone == row 5 of the batch
    one = net.generate(T, inp['x0'][5:6], inp['uniforms'][5:6], lc_up=lc[5:6], gc_ids=inp['gc_ids'][5:6]).cpu().numpy()
    assert np.array_equal(one[0], a[5])
    # oracle on a bounded prefix of two rows
    om = oracle_model(synth.cfg2(2), w)
    so = om.generate(200, inp['x0'][:2], inp['uniforms'][:2, :200], lc_up=lc[:2, :200].cpu().numpy(),
                     gc_ids=inp['gc_ids'][:2], plan=plan_from_dict(net.plan()))
    assert np.array_equal(so, a[:2, :200])
