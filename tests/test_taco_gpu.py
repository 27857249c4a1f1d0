"""GPU parity tests of the Tacotron path: libtaco_b200.so (through the C ABI, via the Tacotron class) against the
numpy oracle on the same seeded inputs and against the committed golden fixtures.  Tolerance: 1e-4 absolute on
float mel / linear / alignment outputs (north_star); observed errors are printed in the assertion messages."""
import os

import numpy as np
import pytest
import torch

from oracle.taco_oracle import TacotronOracle
from tacotron_wavenet_vocoder_korean_b200 import synth
from tacotron_wavenet_vocoder_korean_b200.tacotron import Tacotron
from tests.taco_helpers import Bag, CASES, case, make_batch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TOL = 1e-4


def run_cuda(hp, ns, w, ids, lens, spk, steps, manual=None, want_linear=True):
    m = Tacotron(Bag(hp))
    m.load_state_dict(w)
    if manual is not None:
        m.is_manual_attention, m.manual_alignments = True, manual
    m.initialize(ids, lens, ns, spk, rnn_decoder_test_mode=True, n_steps=steps, want_linear=want_linear)
    torch.cuda.synchronize()
    return m


def check(m, mel, lin, al, tol=TOL):
    e_mel = np.abs(m.mel_outputs.cpu().numpy() - mel).max()
    e_al = np.abs(m.alignments.cpu().numpy() - al).max()
    e_lin = np.abs(m.linear_outputs.cpu().numpy() - lin).max() if m.linear_outputs is not None else 0.0
    assert e_mel <= tol and e_al <= tol and e_lin <= tol, "max |err| mel %.3g linear %.3g alignments %.3g" % (e_mel, e_lin, e_al)
    return e_mel, e_lin, e_al


@pytest.mark.parametrize("name", sorted(CASES))
def test_tiny_cases_match_oracle(name):
    hp, ns, w, ids, lens, spk, steps = case(name)
    mel, lin, al = TacotronOracle(hp, w, ns).synthesize(ids, lens, spk, max_iters=steps)
    m = run_cuda(hp, ns, w, ids, lens, spk, steps)
    check(m, mel, lin, al)
    inf = m.info()
    assert inf['kernel_launches'] > 10 and inf['dec_grid'] == inf['sm_count']


@pytest.mark.parametrize("name", ['tiny_mon_norm', 'tiny_loc_sen'])
def test_matches_golden_fixture(name):
    hp, ns, w, ids, lens, spk, steps = case(name)
    g = np.load(os.path.join(ROOT, 'tests', 'golden', 'taco_%s.npz' % name))
    m = run_cuda(hp, ns, w, g['ids'], g['lens'], spk, steps)
    check(m, g['mel'], g['linear'], g['alignments'])


def test_intermediate_stages_match_oracle():
    hp, ns, w, ids, lens, spk, steps = case('tiny_mon_norm')
    taps = {}
    TacotronOracle(hp, w, ns).synthesize(ids, lens, spk, max_iters=steps, taps=taps)
    m = run_cuda(hp, ns, w, ids, lens, spk, steps)
    pairs = [('enc_prenet', 'enc_prenet'), ('enc_bank', 'encoder_cbhg/bank'), ('enc_highway_in', 'encoder_cbhg/highway_in'),
             ('enc_rnn_in', 'encoder_cbhg/rnn_in'), ('encoder_out', 'encoder_out'), ('post_bank', 'post_cbhg/bank'),
             ('post_highway_in', 'post_cbhg/highway_in'), ('post_rnn_in', 'post_cbhg/rnn_in'), ('post_out', 'post_out')]
    for cname, oname in pairs:
        ref = taps[oname]
        got = m.debug_tensor(cname, ref.shape)
        err = np.abs(got - ref).max()
        assert err <= TOL, "%s: max |err| %.3g" % (cname, err)


def test_host_buffer_entry_point_equals_device_path():
    """taco_synthesize_host (ids in, mel / linear / alignments out through HOST buffers, copies inside): bit-identical to the
    device path, on reused buffers, with and without the linear output; the pinned staging slots are shrunk to 1.25 MB so that
    the 5.7 MB linear output crosses them in several pipelined chunks with a partial last one (and the 1 MB threaded-copy split)."""
    import ctypes as C
    from tacotron_wavenet_vocoder_korean_b200 import _taco_lib
    hp, ns, w, ids, lens, spk, steps = case('tiny_mon_norm')
    reps = 24
    ids, lens, spk = np.tile(ids, (reps, 1)), np.tile(lens, reps), np.tile(spk, reps)
    steps = 160
    m = run_cuda(hp, ns, w, ids, lens, spk, steps)
    os.environ['TACO_STAGE_BYTES'] = str(5 << 18)          # read when the handle first allocates its staging ring
    N, T_in = ids.shape
    r, nm, nf = hp['reduction_factor'], hp['num_mels'], hp['num_freq']
    L = _taco_lib.lib()
    ids_c = np.ascontiguousarray(ids, np.int32)
    lens_c, spk_c = np.ascontiguousarray(lens, np.int32), np.ascontiguousarray(spk, np.int32)
    for want_linear in (True, False, True):
        mel_h = np.full((N, steps * r, nm), np.nan, np.float32)
        lin_h = np.full((N, steps * r, nf), np.nan, np.float32)
        al_h = np.full((N, T_in, steps), np.nan, np.float32)
        a = _taco_lib.TacoSynthArgs()
        a.N, a.T_in, a.n_steps = N, T_in, steps
        a.ids_dev = ids_c.ctypes.data
        a.lengths = lens_c.ctypes.data_as(C.POINTER(C.c_int32))
        a.speaker_ids = spk_c.ctypes.data_as(C.POINTER(C.c_int32))
        a.mel_dev, a.alignments_dev = mel_h.ctypes.data, al_h.ctypes.data
        a.linear_dev = lin_h.ctypes.data if want_linear else None
        assert L.taco_synthesize_host(m._h, C.byref(a)) == 0, L.taco_last_error(m._h)
        assert np.array_equal(mel_h, m.mel_outputs.cpu().numpy())
        assert np.array_equal(al_h, m.alignments.cpu().numpy())
        if want_linear:
            assert np.array_equal(lin_h, m.linear_outputs.cpu().numpy())
        else:
            assert np.isnan(lin_h).all()
    del os.environ['TACO_STAGE_BYTES']


def test_unfused_decoder_plan_agrees_with_fused_plan():
    """TACO_NO_FUSE=1 keeps one phase per dense layer (13 per step); the default plan evaluates the first prenet layer of the next
    step inside the output-projection phase from the product matrix.  Same outputs within fp32 reassociation noise, both within
    1e-4 of the oracle."""
    hp, ns, w, ids, lens, spk, steps = case('tiny_mon_norm')
    mel, lin, al = TacotronOracle(hp, w, ns).synthesize(ids, lens, spk, max_iters=steps)
    fused = run_cuda(hp, ns, w, ids, lens, spk, steps)
    os.environ['TACO_NO_FUSE'] = '1'
    try:
        plain = run_cuda(hp, ns, w, ids, lens, spk, steps)
    finally:
        del os.environ['TACO_NO_FUSE']
    check(fused, mel, lin, al)
    check(plain, mel, lin, al)
    assert plain.info()['dec_phases_per_step'] == fused.info()['dec_phases_per_step'] + 1
    assert np.abs(plain.mel_outputs.cpu().numpy() - fused.mel_outputs.cpu().numpy()).max() < 2e-5


def test_manual_alignments_override():
    hp, ns, w, ids, lens, spk, steps = case('tiny_mon_norm')
    N, T_in = ids.shape
    man = np.zeros((N, steps, T_in), np.float32)
    for t in range(steps):
        man[:, t, min(t, T_in - 1)] = 1.0
    mel, lin, al = TacotronOracle(hp, w, ns).synthesize(ids, lens, spk, manual_alignments=man, max_iters=steps)
    m = run_cuda(hp, ns, w, ids, lens, spk, steps, manual=man)
    check(m, mel, lin, al)


def test_full_size_model_matches_oracle():
    """The reference's hparams.py:124-166 shapes (7.07 M parameters), 4 sentences x 40 tokens x 30 decoder steps."""
    hp = dict(synth.TACO_HP)
    w = synth.make_taco_weights(hp, 2)
    ids, lens, spk = make_batch(4, 40, seed=7)
    steps = 30
    mel, lin, al = TacotronOracle(hp, w, 2).synthesize(ids, lens, spk, max_iters=steps)
    m = run_cuda(hp, 2, w, ids, lens, spk, steps)
    check(m, mel, lin, al)
    inf = m.info()
    assert inf['rnn_weights_in_smem'] == 1 and inf['dec_phases_per_step'] == 12      # output projection + first prenet layer share a phase


def test_full_size_loc_sen_matches_oracle():
    hp = dict(synth.TACO_HP, attention_type='loc_sen')
    w = synth.make_taco_weights(hp, 2)
    ids, lens, spk = make_batch(3, 50, seed=8)
    steps = 20
    mel, lin, al = TacotronOracle(hp, w, 2).synthesize(ids, lens, spk, max_iters=steps)
    m = run_cuda(hp, 2, w, ids, lens, spk, steps)
    check(m, mel, lin, al)


def test_properties_at_full_batch():
    """cfg-3 size (32 sentences, 200 decoder steps): too slow for the numpy oracle, so size-independent properties:
    rows are independent (a sentence synthesised alone, at the same padded length -- the reference's CBHG convolutions
    are not masked, so the padding is part of the input -- gives the same mel as inside the batch), alignments are
    non-negative, vanish past each sentence's length and each step's total mass is <= 1 (monotonic attention),
    and the run is deterministic."""
    hp = dict(synth.TACO_HP)
    w = synth.make_taco_weights(hp, 2)
    ids, lens, spk = make_batch(32, 60, seed=9, min_len=20)
    m = run_cuda(hp, 2, w, ids, lens, spk, 200, want_linear=False)
    mel = m.mel_outputs.cpu().numpy()
    al = m.alignments.cpu().numpy()
    assert mel.shape == (32, 1000, 80) and np.isfinite(mel).all()
    assert (al >= 0).all() and (al.sum(1) <= 1 + 1e-4).all()
    for n in range(32):
        assert np.all(al[n, lens[n]:, :] == 0)
    m2 = run_cuda(hp, 2, w, ids, lens, spk, 200, want_linear=False)
    assert np.array_equal(m2.mel_outputs.cpu().numpy(), mel)
    for n in (0, 5, 31):
        m1 = run_cuda(hp, 2, w, ids[n:n + 1], lens[n:n + 1], spk[n:n + 1], 200, want_linear=False)
        e = np.abs(m1.mel_outputs.cpu().numpy()[0] - mel[n]).max()
        assert e <= TOL, "row %d alone vs in batch: %.3g" % (n, e)


def test_synthesizer_entry_point(tmp_path):
    from synthesizer import Synthesizer
    from tacotron_wavenet_vocoder_korean_b200.hparams import hparams
    w = synth.make_taco_weights({k: getattr(hparams, k) for k in synth.TACO_HP}, 2)
    s = Synthesizer()
    s.load(None, num_speakers=2, weights=w)
    res = s.synthesize(texts=['존경하는 독일 국민 여러분', '고국에 계신 국민 여러분'], base_path=str(tmp_path), speaker_ids=[0, 1])
    assert len(res) == 2
    for r in res:
        assert r['mel'].shape[1] == 80 and r['mel'].shape[0] <= 1000 and os.path.exists(r['mel_path'])
        assert np.array_equal(np.load(r['mel_path']), r['mel'])


def test_end_to_end_text_to_wav_tiny():
    """pipeline.TextToSpeech: tiny Tacotron (20 mels, r=2) feeding the tiny MoL WaveNet (20 lc channels, hop 6);
    the waveform of each sentence must equal the WaveNet oracle run on the mel the Tacotron path produced."""
    import oracle
    from tacotron_wavenet_vocoder_korean_b200 import pipeline
    from tests.helpers import oracle_model, plan_from_dict
    hp = synth.taco_tiny()
    tw = synth.make_taco_weights(hp, 2)
    kw = synth.tiny_mol(batch_size=2)
    ww = synth.make_weights(**kw)
    tts = pipeline.TextToSpeech(Bag(hp), tw, 2, kw, ww, hop_size=6)
    texts = ['안녕하세요', '반갑습니다 여러분', '좋은 아침']
    spk = [0, 1, 1]
    mels = tts.text_to_mel(texts, spk, attention_trim=True)
    assert all(m.shape[1] == 20 and 3 <= m.shape[0] <= 24 for m in mels)
    wavs = tts.synthesize(texts, spk, seed=3)
    assert [len(w) for w in wavs] == [m.shape[0] * 6 for m in mels]
    assert all(np.isfinite(w).all() and np.abs(w).max() <= 1.0 for w in wavs)
    # determinism of the whole chain
    wavs2 = tts.synthesize(texts, spk, seed=3)
    # (uniforms are drawn on the device from torch's generator: only shapes/finite-ness are stable across calls)
    assert [len(w) for w in wavs2] == [len(w) for w in wavs]


@pytest.mark.parametrize("name", sorted(CASES))
def test_tensor_core_conv_path_matches_oracle(name, monkeypatch):
    """The CBHG convolutions / projections / highway / GRU input GEMMs on tcgen05 with the 3xTF32 split (taco_gemm_tc.cuh):
    the tiny cases are below the row threshold of that path, so force it (odd channel counts exercise the zero padding to
    32-channel k-tiles, T < 128 the out-of-bounds rows of the TMA boxes), and compare with the oracle AND with the fp32 SIMT path."""
    hp, ns, w, ids, lens, spk, steps = case(name)
    mel, lin, al = TacotronOracle(hp, w, ns).synthesize(ids, lens, spk, max_iters=steps)
    monkeypatch.setenv('TACO_TC_MIN_ROWS', '1')
    m = run_cuda(hp, ns, w, ids, lens, spk, steps)
    assert m.info()['tc_gemm_launches'] >= 8, m.info()
    check(m, mel, lin, al)
    monkeypatch.setenv('TACO_NO_TC', '1')
    s = run_cuda(hp, ns, w, ids, lens, spk, steps)
    assert s.info()['tc_gemm_launches'] == 0
    for a_, b_ in ((m.mel_outputs, s.mel_outputs), (m.linear_outputs, s.linear_outputs), (m.alignments, s.alignments)):
        assert float((a_ - b_).abs().max()) <= 2e-5


def test_cfg3_tensor_core_path_is_the_default_and_matches_simt(monkeypatch):
    """cfg-3 shape (32 sentences, r = 5): every CBHG GEMM group runs on the tensor cores by default; outputs within 5e-5 of the fp32 path."""
    import sys
    sys.path.insert(0, os.path.join(ROOT, 'scripts'))
    from bench_taco import make_texts
    from tacotron_wavenet_vocoder_korean_b200.text import text_to_sequence, prepare_inputs
    hp = dict(synth.TACO_HP)
    w = synth.make_taco_weights(hp, 2)
    ids = prepare_inputs([text_to_sequence(t) for t in make_texts(32)])
    lens = np.array([int(np.argmax(s == 1)) + 1 for s in ids], np.int32)
    spk = (np.arange(32) % 2).astype(np.int32)
    m = run_cuda(hp, 2, w, ids, lens, spk, 24)
    assert m.info()['tc_gemm_launches'] >= 17, m.info()
    monkeypatch.setenv('TACO_NO_TC', '1')
    s = run_cuda(hp, 2, w, ids, lens, spk, 24)
    assert s.info()['tc_gemm_launches'] == 0
    errs = [float((a_ - b_).abs().max()) for a_, b_ in ((m.mel_outputs, s.mel_outputs), (m.linear_outputs, s.linear_outputs), (m.alignments, s.alignments))]
    assert max(errs) <= 5e-5, errs       # observed 2.4e-5 on mel after 24 recurrent steps: two fp32-accurate summation orders, north_star allows 1e-4


@pytest.mark.parametrize("att,steps,n,t_in", [('bah_mon_norm', 200, 4, 60), ('loc_sen', 120, 3, 50)])
def test_full_size_model_matches_oracle_over_the_whole_decode(att, steps, n, t_in):
    """cfg-3's 200 decoder steps (1000 mel frames) against the numpy oracle on the full-size model, and the post-CBHG /
    linear outputs over all of them: the 30- / 20-step tests above stop before the attention has walked through the sentence."""
    hp = dict(synth.TACO_HP, attention_type=att)
    w = synth.make_taco_weights(hp, 2)
    ids, lens, spk = make_batch(n, t_in, seed=11, min_len=t_in // 2)
    mel, lin, al = TacotronOracle(hp, w, 2).synthesize(ids, lens, spk, max_iters=steps)
    m = run_cuda(hp, 2, w, ids, lens, spk, steps)
    e = check(m, mel, lin, al)
    assert m.mel_outputs.shape[1] == steps * hp['reduction_factor'], e
