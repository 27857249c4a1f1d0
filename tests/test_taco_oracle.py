"""CPU tests of the Tacotron path: the numpy oracle is pinned against torch.nn.functional for the TF ops whose
'same' semantics it restates, against the recursive definition of monotonic attention, against an fp64
evaluation and against the committed golden fixtures; the host logic (tokeniser, trim, config validation) and
the C ABI symbol table are checked without a GPU."""
import ctypes as C
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import taco_oracle as to
from tacotron_wavenet_vocoder_korean_b200 import _taco_lib, synth
from tests.taco_helpers import Bag, case, make_batch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("k", [1, 2, 3, 4, 7, 16, 31])
def test_conv1d_same_matches_torch(k):
    rs = np.random.RandomState(k)
    x = rs.randn(2, 13, 5).astype(np.float32)
    W = rs.randn(k, 5, 6).astype(np.float32)
    b = rs.randn(6).astype(np.float32)
    y = to.conv1d_same(x, W, b)
    pl, pr = (k - 1) // 2, (k - 1) - (k - 1) // 2             # TF 'SAME': extra padding goes to the right
    xt = F.pad(torch.from_numpy(x).transpose(1, 2), (pl, pr))
    yt = F.conv1d(xt, torch.from_numpy(W).permute(2, 1, 0), torch.from_numpy(b)).transpose(1, 2).numpy()
    assert np.allclose(y, yt, atol=1e-5)


def test_maxpool_and_batchnorm_match_torch():
    rs = np.random.RandomState(0)
    x = rs.randn(3, 9, 4).astype(np.float32)
    yt = F.max_pool1d(F.pad(torch.from_numpy(x).transpose(1, 2), (0, 1), value=float('-inf')), 2, 1).transpose(1, 2).numpy()
    assert np.array_equal(to.maxpool2_same(x), yt)
    g, b, m, v = (rs.rand(4).astype(np.float32) + 0.5 for _ in range(4))
    bt = F.batch_norm(torch.from_numpy(x).transpose(1, 2), torch.from_numpy(m), torch.from_numpy(v), torch.from_numpy(g),
                      torch.from_numpy(b), False, 0.0, 1e-3).transpose(1, 2).numpy()
    assert np.allclose(to.batch_norm(x, g, b, m, v), bt, atol=1e-6)


def test_gru_cell_definition():
    """tf.contrib.rnn.GRUCell: r applied to the state BEFORE the candidate matmul (unlike torch.nn.GRUCell)."""
    rs = np.random.RandomState(1)
    n_in, U = 5, 4
    x, h = rs.randn(n_in), rs.randn(U)
    Wg, bg, Wc, bc = rs.randn(n_in + U, 2 * U), rs.randn(2 * U), rs.randn(n_in + U, U), rs.randn(U)
    out = to.gru_cell(x, h, Wg, bg, Wc, bc)
    r = np.zeros(U)
    u = np.zeros(U)
    for c in range(U):
        r[c] = 1 / (1 + np.exp(-(sum(x[i] * Wg[i, c] for i in range(n_in)) + sum(h[i] * Wg[n_in + i, c] for i in range(U)) + bg[c])))
        u[c] = 1 / (1 + np.exp(-(sum(x[i] * Wg[i, U + c] for i in range(n_in)) + sum(h[i] * Wg[n_in + i, U + c] for i in range(U)) + bg[U + c])))
    ref = np.zeros(U)
    for c in range(U):
        cand = np.tanh(sum(x[i] * Wc[i, c] for i in range(n_in)) + sum(r[i] * h[i] * Wc[n_in + i, c] for i in range(U)) + bc[c])
        ref[c] = u[c] * h[c] + (1 - u[c]) * cand
    assert np.allclose(out, ref, atol=1e-12)


def test_monotonic_parallel_matches_recursive_definition():
    rs = np.random.RandomState(2)
    p = rs.uniform(0.02, 0.6, (4, 20))                          # keeps cumprod(1-p) above the 1e-10 clip of the closed form
    p[1, 12:] = 0.0                                            # masked tail
    prev = np.zeros((4, 20))
    prev[:, 0] = 1.0
    for _ in range(6):
        a = to.monotonic_attention_parallel(p, prev)
        b = to.monotonic_attention_recursive(p, prev)
        assert np.allclose(a, b, atol=1e-12)
        assert np.all(a.sum(1) <= 1 + 1e-9)
        prev = a


def test_birnn_batched_equals_per_row():
    hp, ns, w, ids, lens, spk, steps = case('tiny_mon_norm')
    o = to.TacotronOracle(hp, w, ns)
    rs = np.random.RandomState(3)
    x = rs.randn(3, 11, 16).astype(np.float32)
    init = rs.randn(3, 32).astype(np.float32)
    a = o._birnn(x, lens, 'encoder_cbhg', 16, init[:, :16], init[:, 16:])
    b = o._birnn_batched(x, lens, 'encoder_cbhg', 16, init[:, :16], init[:, 16:])
    assert np.allclose(a, b, atol=1e-6)
    assert np.all(a[1, lens[1]:] == 0)


@pytest.mark.parametrize("name", ['tiny_mon_norm', 'tiny_mon', 'tiny_loc_sen', 'tiny_single_speaker'])
def test_fp32_oracle_close_to_fp64(name):
    hp, ns, w, ids, lens, spk, steps = case(name)
    m32, l32, a32 = to.TacotronOracle(hp, w, ns, np.float32).synthesize(ids, lens, spk, max_iters=steps)
    m64, l64, a64 = to.TacotronOracle(hp, w, ns, np.float64).synthesize(ids, lens, spk, max_iters=steps)
    assert m32.shape == (ids.shape[0], steps * hp['reduction_factor'], hp['num_mels'])
    assert l32.shape == (ids.shape[0], steps * hp['reduction_factor'], hp['num_freq'])
    assert a32.shape == (ids.shape[0], ids.shape[1], steps)
    assert np.abs(m32 - m64).max() < 1e-5 and np.abs(l32 - l64).max() < 1e-5 and np.abs(a32 - a64).max() < 1e-5
    assert np.all(a32[1, lens[1]:, :] == 0), "alignment past the sentence length must be exactly zero"
    assert np.abs(m32).max() > 1e-2


@pytest.mark.parametrize("name", ['tiny_mon_norm', 'tiny_loc_sen'])
def test_oracle_matches_golden(name):
    hp, ns, w, ids, lens, spk, steps = case(name)
    g = np.load(os.path.join(ROOT, 'tests', 'golden', 'taco_%s.npz' % name))
    assert np.array_equal(g['ids'], ids) and np.array_equal(g['lens'], lens)
    mel, lin, al = to.TacotronOracle(hp, w, ns).synthesize(ids, lens, spk, max_iters=steps)
    assert np.allclose(mel, g['mel'], atol=1e-6) and np.allclose(lin, g['linear'], atol=1e-6) and np.allclose(al, g['alignments'], atol=1e-6)


def test_manual_alignment_override():
    hp, ns, w, ids, lens, spk, steps = case('tiny_mon_norm')
    N, T_in = ids.shape
    man = np.zeros((N, steps, T_in), np.float32)
    for t in range(steps):
        man[:, t, min(t, T_in - 1)] = 1.0
    mel, lin, al = to.TacotronOracle(hp, w, ns).synthesize(ids, lens, spk, manual_alignments=man, max_iters=steps)
    assert np.array_equal(al, man.transpose(0, 2, 1))


def test_full_size_shapes_and_param_count():
    hp = dict(synth.TACO_HP)
    shapes = synth.taco_weight_shapes(hp, 2)
    n = sum(int(np.prod(s)) for s in shapes.values())
    assert n == 7069763
    # per-step decoder weights (about 1.63 M MAC per row and step)
    dec = sum(int(np.prod(s)) for k, s in shapes.items() if '/decoder/' in k and k.endswith('kernel'))
    assert 1.60e6 < dec < 1.66e6


# ---- host logic ---------------------------------------------------------------------------------------
def test_tokeniser_known_answers():
    from text import text_to_sequence, sequence_to_text
    # text/__init__.py:70 comment of the reference: '존경하는' -> [14, 29, 45, 2, 27, 62, 20, 21, 4, 39, 45, 1]
    assert list(text_to_sequence('존경하는')) == [14, 29, 45, 2, 27, 62, 20, 21, 4, 39, 45, 1]
    from tacotron_wavenet_vocoder_korean_b200.text.korean import ALL_SYMBOLS, char_to_id, normalize
    assert len(ALL_SYMBOLS) == 80 and char_to_id['ᄀ'] == 2 and char_to_id['ᅡ'] == 21 and char_to_id['ᆨ'] == 42 and char_to_id[' '] == 79
    s = '고국에 계신 국민 여러분, 안녕하십니까?'
    assert sequence_to_text(text_to_sequence(s), skip_eos_and_pad=True, combine_jamo=True) == s
    assert normalize('오늘(13일) 60.3%') == '오늘 육십쩜 삼퍼센트'
    assert normalize('19가지와 JTBC') == '열아홉가지와 제이티비씨'
    assert normalize('2017년 9월 12일') == '이천일십칠년 구월 십이일'


def test_prepare_inputs_and_lengths():
    from text import text_to_sequence, prepare_inputs
    seqs = prepare_inputs([text_to_sequence('가'), text_to_sequence('가나다')])
    assert seqs.shape == (2, 7) and seqs[0, 3:].tolist() == [0] * 4
    assert [int(np.argmax(a == 1)) + 1 for a in seqs] == [3, 7]


def test_attention_trim_index():
    from synthesizer import attention_trim_index
    al = np.zeros((6, 10), np.float32)
    path = [0, 1, 2, 3, 4, 5, 5, 5, 5, 5]
    for t, j in enumerate(path):
        al[j, t] = 1
    # last token (5) first attended at step 5 and held: stops after it has been seen min(count, 5) times
    assert attention_trim_index(al, 6, 5) == 5 * 9 + 3
    al2 = np.zeros((6, 4), np.float32)
    for t, j in enumerate([0, 2, 3, 1]):
        al2[j, t] = 1
    assert attention_trim_index(al2, 6, 5) == 5 * 2 + 3


# ---- C ABI (no compute without a GPU) -----------------------------------------------------------------
def test_taco_cabi_exports_every_declared_symbol():
    import re
    hdr = open(os.path.join(ROOT, 'include', 'taco_b200.h')).read()
    declared = set(re.findall(r'\b(taco_[a-z_]+)\s*\(', hdr))
    assert declared == set(_taco_lib.EXPORTS)
    L = _taco_lib.lib()
    for name in declared:
        assert hasattr(L, name), name


def test_taco_config_struct_matches_header_and_validation():
    L = _taco_lib.lib()
    hp = Bag(synth.TACO_HP)
    cfg = _taco_lib.make_config(hp, 2)
    assert C.sizeof(cfg) == 4 * (4 + 5 + 4 + 5 + 1 + 3 + 2 + 5 + 4 + 6 + 4)
    h = C.c_void_p()
    assert L.taco_create(C.byref(cfg), C.byref(h)) == 0
    # compute entry points must fail loudly, not fall back, when weights / device are missing
    assert L.taco_finalize(h) != 0
    assert L.taco_last_error(h)
    a = _taco_lib.TacoSynthArgs()
    assert L.taco_synthesize(h, C.byref(a), None) == -2
    L.taco_destroy(h)
    bad = _taco_lib.make_config(hp, 2)
    bad.attention_size = 100
    assert L.taco_create(C.byref(bad), C.byref(h)) == -1 and b'multiple of 32' in L.taco_last_error(None)
    with pytest.raises(NotImplementedError):
        _taco_lib.make_config(dict(synth.TACO_HP, attention_type='luong'), 1)


def test_tacotron_class_surface():
    from tacotron import Tacotron, create_model
    m = create_model(Bag(synth.TACO_HP))
    assert isinstance(m, Tacotron)
    with pytest.raises(NotImplementedError):
        m.initialize(np.zeros((1, 3), np.int32), [3], 1, None, rnn_decoder_test_mode=False)
    with pytest.raises(NotImplementedError):
        m.add_loss()
    if not torch.cuda.is_available():
        m.load_state_dict({})
        with pytest.raises(RuntimeError):
            m.initialize(np.zeros((1, 3), np.int32), [3], 1, None, rnn_decoder_test_mode=True)
