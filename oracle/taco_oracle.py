"""TEST INFRASTRUCTURE -- numpy restatement of the reference's Tacotron inference graph (SURVEY.md
rows a15-a20, Appendix C).  Only tests/, __graft_entry__.smoke() and the CPU legs of the benchmarks may
import it; the product package never does.

PARITY: pinned to the reference's OWN tacotron package.  tacotron/tacotron.py (Tacotron.initialize, inference), rnn_wrappers.py
(AttentionWrapper incl. the manual-alignment override, DecoderPrenetWrapper, ConcatOutputAndAttentionWrapper,
LocationSensitiveAttention), helpers.py (TacoTestHelper) and modules.py are imported unmodified from /root/reference and run on
numpy stand-ins for the TF API (tests/golden/tf_numpy_shim.py, tf_contrib_shim.py); this oracle reproduces the resulting mel /
linear / alignment outputs to 2.4e-7 for bah_mon_norm, bah_mon, loc_sen, one speaker, the post-net dense branch and manual
alignments (tests/golden/make_reference_taco_full_golden.py -> ref_taco_full_*.npz, tests/test_reference_pin.py).
UNPINNED: the arithmetic of the tf.contrib / tf.layers classes themselves (GRUCell, BahdanauMonotonicAttention incl.
monotonic_attention, dynamic_decode, conv1d / batch_normalization / max_pooling1d 'same', bidirectional_dynamic_rnn with
sequence_length), restated here and in the stand-in from their published definitions; TensorFlow 1.x is not installable.  Local
pins of those pieces (tests/test_taco_oracle.py): conv1d-'same' / max-pool / batch-norm against torch.nn.functional, the parallel
monotonic-attention closed form against the recursive definition of Raffel et al. 2017, fp32 against fp64 evaluation.

Follows (reference file:line):
  tacotron/tacotron.py:36-235      graph wiring (embedding zero row :51-56, deepvoice speaker states :78-84,
                                   encoder :103-112, attention cell :128-152, decoder stack :165-201, post :204-219)
  tacotron/modules.py:15-23        prenet;   :25-74 cbhg;   :83-89 highwaynet;   :92-96 conv1d + batch norm
  tacotron/rnn_wrappers.py:282-398 AttentionWrapper.call / _compute_attention (manual alignment override :374)
  tacotron/rnn_wrappers.py:423-430 DecoderPrenetWrapper;  :457-464 ConcatOutputAndAttentionWrapper
  tacotron/rnn_wrappers.py:647-726 LocationSensitiveAttention ('loc_sen')
  tacotron/helpers.py:29-41        TacoTestHelper (go frame, feed last of r frames, stop on all-zero output)
"""
import numpy as np


def sigmoid(x):
    return 1.0 / (1.0 + np.exp(-x))


def softsign(x):
    return x / (np.abs(x) + 1)


def dense(x, W, b=None):
    y = x @ W
    return y if b is None else y + b


def gru_cell(x, h, Wg, bg, Wc, bc):
    """tf.contrib.rnn.GRUCell.call: gates = sigmoid([x,h] Wg + bg) -> r,u; c = tanh([x, r*h] Wc + bc);
    h' = u*h + (1-u)*c."""
    n = h.shape[-1]
    g = sigmoid(np.concatenate([x, h], -1) @ Wg + bg)
    r, u = g[..., :n], g[..., n:]
    c = np.tanh(np.concatenate([x, r * h], -1) @ Wc + bc)
    return u * h + (1 - u) * c


def conv1d_same(x, W, b):
    """tf.layers.conv1d(padding='same', strides=1): x (N,T,Ci), W (k,Ci,Co) cross-correlation,
    pad_left = (k-1)//2, pad_right = k-1-pad_left."""
    k = W.shape[0]
    N, T, _ = x.shape
    pl = (k - 1) // 2
    xp = np.zeros((N, T + k - 1, x.shape[2]), x.dtype)
    xp[:, pl:pl + T] = x
    y = np.zeros((N, T, W.shape[2]), x.dtype)
    for j in range(k):
        y += xp[:, j:j + T] @ W[j]
    return y + b


def batch_norm(x, gamma, beta, mean, var, eps=1e-3):
    """tf.layers.batch_normalization(training=False), epsilon default 1e-3."""
    inv = gamma / np.sqrt(var + x.dtype.type(eps))
    return x * inv + (beta - mean * inv)


def maxpool2_same(x):
    """tf.layers.max_pooling1d(pool_size=2, strides=1, padding='same'): pads one step on the right."""
    y = x.copy()
    y[:, :-1] = np.maximum(x[:, :-1], x[:, 1:])
    return y


def monotonic_attention_parallel(p, prev):
    """tf.contrib.seq2seq.monotonic_attention(mode='parallel')."""
    dt = p.dtype
    tiny = np.finfo(dt).tiny
    one_m = np.clip(1 - p, tiny, 1)
    ex = np.concatenate([np.zeros_like(p[:, :1]), np.cumsum(np.log(one_m), axis=1)[:, :-1]], axis=1)  # exclusive cumsum
    cp = np.exp(ex)
    return p * cp * np.cumsum(prev / np.clip(cp, dt.type(1e-10), 1), axis=1)


def monotonic_attention_recursive(p, prev):
    """Definition (Raffel et al. 2017, eq. 10): a_j = p_j * ((1-p_{j-1}) a_{j-1} / p_{j-1} + prev_j),
    evaluated through q_j = (1-p_{j-1}) q_{j-1} + prev_j, a_j = p_j q_j (no division)."""
    q = np.zeros_like(p[:, 0])
    out = np.zeros_like(p)
    for j in range(p.shape[1]):
        q = (1 - p[:, j - 1]) * q + prev[:, j] if j > 0 else prev[:, 0]
        out[:, j] = p[:, j] * q
    return out


DEFAULT_HP = dict(
    num_symbols=80, embedding_size=256, speaker_embedding_size=16, model_type='deepvoice',
    enc_prenet_sizes=[256, 128], enc_bank_size=16, enc_bank_channel_size=128, enc_maxpool_width=2,
    enc_highway_depth=4, enc_rnn_size=128, enc_proj_sizes=[128, 128], enc_proj_width=3,
    attention_type='bah_mon_norm', attention_size=256, attention_state_size=256,
    dec_layer_num=2, dec_rnn_size=256, dec_prenet_sizes=[256, 128],
    post_bank_size=8, post_bank_channel_size=128, post_maxpool_width=2, post_highway_depth=4,
    post_rnn_size=128, post_proj_sizes=[256, 80], post_proj_width=3,
    reduction_factor=5, max_iters=200, num_mels=80, num_freq=1025,
)


class TacotronOracle(object):
    def __init__(self, hp, weights, num_speakers, dtype=np.float32):
        self.hp = dict(DEFAULT_HP)
        self.hp.update(hp)
        self.dt = np.dtype(dtype)
        self.w = {k: np.asarray(v, dtype=self.dt) for k, v in weights.items()}
        self.num_speakers = num_speakers

    # ---- building blocks -------------------------------------------------------------------
    def W(self, name):
        return self.w['model/inference/' + name]

    def _dense(self, x, name, bias=True):
        return dense(x, self.W(name + '/kernel'), self.W(name + '/bias') if bias else None)

    def _conv_bn(self, x, scope, act):
        y = conv1d_same(x, self.W(scope + '/conv1d/kernel'), self.W(scope + '/conv1d/bias'))
        if act:
            y = np.maximum(y, 0)
        bn = scope + '/batch_normalization/'
        return batch_norm(y, self.W(bn + 'gamma'), self.W(bn + 'beta'), self.W(bn + 'moving_mean'), self.W(bn + 'moving_variance'))

    def _gru(self, x, h, scope):
        return gru_cell(x, h, self.W(scope + '/gates/kernel'), self.W(scope + '/gates/bias'),
                        self.W(scope + '/candidate/kernel'), self.W(scope + '/candidate/bias'))

    def _birnn(self, x, lengths, scope, rnn_size, init_fw, init_bw):
        """tf.nn.bidirectional_dynamic_rnn(sequence_length=lengths): past the length the output is zero and
        the state is carried; the backward direction runs over the length-reversed sequence."""
        N, T, _ = x.shape
        if lengths is None:
            lengths = np.full((N,), T, np.int64)
        out = np.zeros((N, T, 2 * rnn_size), self.dt)
        for n in range(N):
            L = int(lengths[n])
            h = init_fw[n] if init_fw is not None else np.zeros(rnn_size, self.dt)
            for t in range(L):
                h = self._gru(x[n, t], h, scope + '/bidirectional_rnn/fw/gru_cell')
                out[n, t, :rnn_size] = h
            h = init_bw[n] if init_bw is not None else np.zeros(rnn_size, self.dt)
            for t in range(L - 1, -1, -1):
                h = self._gru(x[n, t], h, scope + '/bidirectional_rnn/bw/gru_cell')
                out[n, t, rnn_size:] = h
        return out

    def _birnn_batched(self, x, lengths, scope, rnn_size, init_fw, init_bw):
        """Same as _birnn, vectorised over the batch (identical per-row arithmetic)."""
        N, T, _ = x.shape
        if lengths is None:
            lengths = np.full((N,), T, np.int64)
        lengths = np.asarray(lengths)
        out = np.zeros((N, T, 2 * rnn_size), self.dt)
        h = init_fw.copy() if init_fw is not None else np.zeros((N, rnn_size), self.dt)
        for t in range(T):
            live = t < lengths
            if not live.any():
                break
            hn = self._gru(x[:, t], h, scope + '/bidirectional_rnn/fw/gru_cell')
            h = np.where(live[:, None], hn, h)
            out[:, t, :rnn_size] = np.where(live[:, None], hn, 0)
        h = init_bw.copy() if init_bw is not None else np.zeros((N, rnn_size), self.dt)
        for t in range(T - 1, -1, -1):
            live = t < lengths
            if not live.any():
                continue
            hn = self._gru(x[:, t], h, scope + '/bidirectional_rnn/bw/gru_cell')
            h = np.where(live[:, None], hn, h)
            out[:, t, rnn_size:] = np.where(live[:, None], hn, 0)
        return out

    def cbhg(self, x, lengths, scope, K, proj_sizes, depth, rnn_size, before_highway=None, rnn_init=None, taps=None):
        bank = np.concatenate([self._conv_bn(x, '%s/conv_bank/conv1d_%d' % (scope, k), True) for k in range(1, K + 1)], -1)
        if taps is not None:
            taps[scope + '/bank'] = bank
        y = maxpool2_same(bank)
        for i, _ in enumerate(proj_sizes):
            y = self._conv_bn(y, '%s/proj_%d' % (scope, i + 1), i != len(proj_sizes) - 1)
        y = y + x
        if before_highway is not None:
            y = y + before_highway[:, None, :]
        if y.shape[2] != rnn_size:
            y = self._dense(y, scope + '/dense')
        if taps is not None:
            taps[scope + '/highway_in'] = y
        for i in range(depth):
            H = np.maximum(self._dense(y, '%s/highway_%d/H' % (scope, i + 1)), 0)
            Tg = sigmoid(self._dense(y, '%s/highway_%d/T' % (scope, i + 1)))
            y = H * Tg + y * (1 - Tg)
        if taps is not None:
            taps[scope + '/rnn_in'] = y
        fw = bw = None
        if rnn_init is not None:
            fw, bw = rnn_init[:, :rnn_size], rnn_init[:, rnn_size:]
        return self._birnn_batched(y, lengths, scope, rnn_size, fw, bw)

    # ---- the graph ---------------------------------------------------------------------------
    def encode(self, ids, lengths, speaker_ids, taps=None):
        hp = self.hp
        table = self.W('embedding').copy()
        table[0] = 0                                                     # tacotron.py:51-56
        x = table[ids]
        N = ids.shape[0]
        st = {}
        if self.num_speakers > 1:
            assert hp['model_type'] == 'deepvoice' and hp['speaker_embedding_size'] != 1
            se = self.W('speaker_embedding')[speaker_ids]
            names = ['dense', 'dense_1', 'dense_2'] + ['dense_%d' % (3 + i) for i in range(hp['dec_layer_num'])]
            vals = [softsign(self._dense(se, n)) for n in names]          # tacotron.py:76-84
            st['before_highway'], st['enc_init'], st['att_init'] = vals[:3]
            st['dec_init'] = vals[3:]
        else:
            st['before_highway'] = st['enc_init'] = None
            st['att_init'] = np.zeros((N, hp['attention_state_size']), self.dt)
            st['dec_init'] = [np.zeros((N, hp['dec_rnn_size']), self.dt) for _ in range(hp['dec_layer_num'])]
        p = x
        for i, _ in enumerate(hp['enc_prenet_sizes']):
            p = np.maximum(self._dense(p, 'prenet/dense_%d' % (i + 1)), 0)
        if taps is not None:
            taps['enc_prenet'] = p
        enc = self.cbhg(p, lengths, 'encoder_cbhg', hp['enc_bank_size'], hp['enc_proj_sizes'], hp['enc_highway_depth'],
                        hp['enc_rnn_size'], st['before_highway'], st['enc_init'], taps=taps)
        return enc, st

    def decode(self, memory, lengths, st, manual_alignments=None, max_iters=None):
        hp = self.hp
        dt = self.dt
        N, T_in, _ = memory.shape
        lengths = np.asarray(lengths)
        mask = np.arange(T_in)[None, :] < lengths[:, None]
        values = memory * mask[:, :, None]                                # _prepare_memory
        keys = values @ self.W('memory_layer/kernel')
        att = hp['attention_type']
        D = 'decoder/'
        if att == 'bah_mon_norm' or att == 'bah_mon':
            v = self.W(D + 'attention/attention_v')
            if att == 'bah_mon_norm':
                nv = self.W(D + 'attention/attention_g') * v / np.sqrt(np.sum(np.square(v)))
                ab = self.W(D + 'attention/attention_b')
            else:
                nv, ab = v, np.zeros_like(v)
            sb = self.W(D + 'attention/attention_score_bias')
            state = np.zeros((N, T_in), dt)
            state[:, 0] = 1                                               # dirac initial alignments
        elif att == 'loc_sen':
            nv = self.W(D + 'attention/attention_variable')
            ab = self.W(D + 'attention/attention_bias')
            Wcv = self.W(D + 'attention/location_features_convolution/kernel')      # (31,1,32)
            bcv = self.W(D + 'attention/location_features_convolution/bias')
            Wl = self.W(D + 'attention/location_features_layer/kernel')             # (32,units)
            state = np.zeros((N, T_in), dt)
        else:
            raise ValueError(att)
        h_att = st['att_init'].copy()
        hs = [h.copy() for h in st['dec_init']]
        ctx = np.zeros((N, memory.shape[2]), dt)
        r = hp['reduction_factor']
        nm = hp['num_mels']
        x = np.zeros((N, nm), dt)
        outs, aligns = [], []
        finished = np.zeros((N,), bool)
        iters = max_iters or hp['max_iters']
        for t in range(iters):
            p = x
            for i, _ in enumerate(hp['dec_prenet_sizes']):
                p = np.maximum(self._dense(p, D + 'decoder_prenet/dense_%d' % (i + 1)), 0)
            h_att = self._gru(np.concatenate([p, ctx], -1), h_att, D + 'attention_cell/gru_cell')
            q = h_att @ self.W(D + 'attention/query_layer/kernel')
            if att == 'loc_sen':
                f = conv1d_same(state[:, :, None], Wcv, bcv)
                loc = f @ Wl
                e = np.sum(nv * np.tanh(keys + q[:, None, :] + loc + ab), -1)
                e = np.where(mask, e, -np.inf)
                e = e - e.max(-1, keepdims=True)
                ex = np.exp(e)
                al = ex / ex.sum(-1, keepdims=True)
                state = al + state
            else:
                score = np.sum(nv * np.tanh(keys + q[:, None, :] + ab), -1) + sb
                with np.errstate(over='ignore'):
                    pc = np.where(mask, sigmoid(np.where(mask, score, 0)), 0).astype(dt)
                al = monotonic_attention_parallel(pc, state)
                state = al
            used = al if manual_alignments is None else manual_alignments[:, t, :].astype(dt)
            ctx = np.einsum('nt,ntc->nc', used, values).astype(dt)
            aligns.append(used)
            o = self._dense(np.concatenate([h_att, ctx], -1), D + 'concat_projection')
            for i in range(hp['dec_layer_num']):
                hs[i] = self._gru(o, hs[i], D + 'cell_%d/gru_cell' % (i + 1))
                o = o + hs[i]
            out = self._dense(o, D + 'output_projection')
            outs.append(out)
            x = out[:, -nm:]
            finished = finished | np.all(out == 0, axis=1)                # helpers.py:38
            if finished.all():
                break
        dec = np.stack(outs, 1)                                           # (N, steps, nm*r)
        mel = dec.reshape(N, -1, nm)
        alignments = np.stack(aligns, 2)                                  # (N, T_in, steps)
        return mel, alignments

    def post(self, mel, taps=None):
        hp = self.hp
        y = self.cbhg(mel, None, 'post_cbhg', hp['post_bank_size'], hp['post_proj_sizes'], hp['post_highway_depth'],
                      hp['post_rnn_size'], taps=taps)
        if taps is not None:
            taps['post_out'] = y
        return self._dense(y, 'dense_%d' % (3 + hp['dec_layer_num']) if self.num_speakers > 1 else 'dense')

    def synthesize(self, ids, lengths, speaker_ids=None, manual_alignments=None, max_iters=None, taps=None):
        ids = np.asarray(ids)
        if speaker_ids is None:
            speaker_ids = np.zeros((ids.shape[0],), np.int64)
        enc, st = self.encode(ids, lengths, np.asarray(speaker_ids), taps=taps)
        if taps is not None:
            taps['encoder_out'] = enc
        mel, al = self.decode(enc, lengths, st, manual_alignments, max_iters)
        lin = self.post(mel, taps=taps)
        return mel, lin, al
