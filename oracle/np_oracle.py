"""TEST INFRASTRUCTURE -- independent numpy restatement of the reference's WaveNet
incremental generation, written to mirror the *structure* of the reference code
(one call per audio sample, queues shifted by full copies, host categorical draw).

It serves two purposes:
  1. an independent check of oracle/wn_oracle.c (natural plan): same algorithm, but
     libm/numpy transcendentals and BLAS summation order -> agreement to ~1e-5;
  2. the "reference-structure" CPU baseline (BASELINE.md section 3, B-ref).

PARITY: see wn_oracle.c header (pinned to the reference's own Python on a numpy TF stand-in; TF kernels unpinned).

Reference lines followed: wavenet/model.py:41-46,49-64,66-101,102-111,112-167,181-212,
215-245; wavenet/mixture.py:84-114; generate.py:184-233; wavenet/ops.py:22-47.
"""
import numpy as np

F32 = np.float32


def calculate_receptive_field(filter_width, dilations, scalar_input, initial_filter_width):
    # wavenet/model.py:31-39
    rf = (filter_width - 1) * sum(dilations) + 1
    rf += (initial_filter_width - 1) if scalar_input else (filter_width - 1)
    return rf


def mu_law_encode(audio, quantization_channels):
    # wavenet/ops.py:22-33
    audio = np.asarray(audio, F32)
    mu = F32(quantization_channels - 1)
    safe = np.minimum(np.abs(audio), F32(1.0))
    magnitude = np.log1p(mu * safe) / np.log1p(mu)
    signal = np.sign(audio) * magnitude
    return ((signal + 1) / 2 * mu + F32(0.5)).astype(np.int32)


def mu_law_decode(output, quantization_channels, quantization=True):
    # wavenet/ops.py:36-47
    mu = quantization_channels - 1
    if quantization:
        signal = 2 * (np.asarray(output, F32) / F32(mu)) - 1
    else:
        signal = np.asarray(output, F32)
    magnitude = F32(1 / mu) * (F32(1 + mu) ** np.abs(signal) - 1)
    return (np.sign(signal) * magnitude).astype(F32)


def _sigmoid(x):
    return F32(1) / (F32(1) + np.exp(-x))


class NumpyWaveNet:
    """Holds TF-named weights; `step` is predict_proba_incremental (model.py:215-245)."""

    def __init__(self, batch_size, dilations, filter_width, residual_channels, dilation_channels,
                 skip_channels, quantization_channels=256, out_channels=30, use_biases=False,
                 scalar_input=False, initial_filter_width=32, global_condition_channels=None,
                 global_condition_cardinality=None, local_condition_channels=80, upsample_factor=None,
                 train_mode=False):
        assert filter_width == 2
        self.N = batch_size
        self.dilations = list(dilations)
        self.R, self.D, self.S = residual_channels, dilation_channels, skip_channels
        self.Q, self.O = quantization_channels, out_channels
        self.use_biases, self.scalar_input, self.ifw = use_biases, scalar_input, initial_filter_width
        self.G, self.card, self.C = global_condition_channels, global_condition_cardinality, local_condition_channels
        self.upsample_factor = list(upsample_factor or [])
        self.receptive_field = calculate_receptive_field(2, self.dilations, scalar_input, initial_filter_width)
        self.w = {}
        self.reset_queues()

    def set_weights(self, state):
        self.w = {k: np.asarray(v, F32) for k, v in state.items()}

    def reset_queues(self):
        # model.py:49-64 (queue_initializer): all zeros, index 0 = oldest
        N = self.N
        if self.scalar_input:
            self.causal_queue = np.zeros((N, self.ifw, 1), F32)
        else:
            self.causal_queue = np.zeros((N, 2, self.Q), F32)
        if self.C:
            self.lc_queue = np.zeros((N, 2, self.C), F32)
        self.dq = [np.zeros((N, d + 1, self.R), F32) for d in self.dilations]

    def _b(self, name, n):
        return self.w.get(name, np.zeros(n, F32))

    def create_upsample(self, mel):
        # model.py:102-111; conv2d_transpose(kernel (F,2), strides (F,1), 'same'):
        #   out[i*F+a, w] = in[i,w]*K[a,0] + in[i,w-1]*K[a,1]
        x = np.asarray(mel, F32)
        for i, Fk in enumerate(self.upsample_factor):
            K = self.w['wavenet/upsample%d/kernel' % i].reshape(Fk, 2)
            shifted = np.concatenate([np.zeros_like(x[:, :, :1]), x[:, :, :-1]], axis=2)
            y = x[:, :, None, :] * K[None, None, :, 0, None] + shifted[:, :, None, :] * K[None, None, :, 1, None]
            x = y.reshape(x.shape[0], x.shape[1] * Fk, x.shape[2]).astype(F32)
        return x

    def step(self, x_in, lc_row=None, gc_ids=None):
        """x_in: (N,) float (scalar) or int ids.  Returns conv2 output (N, O|Q)."""
        N = self.N
        if self.scalar_input:
            enc = np.asarray(x_in, F32).reshape(N, 1, 1)
        else:
            enc = np.zeros((N, 1, self.Q), F32)
            enc[np.arange(N), 0, np.asarray(x_in, np.int64)] = 1
        gvec = None
        if self.G:
            gvec = self.w['wavenet/gc_embedding'][np.asarray(gc_ids)]          # model.py:194-195
        # queue updates are full shift-copies, as tf.scatter_update(tf.concat(...)) (model.py:122,125)
        self.causal_queue = np.concatenate([self.causal_queue[:, 1:], enc], axis=1)
        if self.C:
            self.lc_queue = np.concatenate([self.lc_queue[:, 1:], np.asarray(lc_row, F32).reshape(N, 1, self.C)], axis=1)
        wc = self.w['wavenet/conv1d/kernel']
        cur = np.einsum('nki,kir->nr', self.causal_queue, wc).astype(F32)   # model.py:41-46 (valid conv, no bias)
        total = None
        for l, d in enumerate(self.dilations):
            p = 'wavenet/dilated_stack/layer%d/dilation_layer/' % l
            self.dq[l] = np.concatenate([self.dq[l][:, 1:], cur[:, None, :]], axis=1)   # model.py:145
            old, now = self.dq[l][:, 0], self.dq[l][:, d]
            wf, wg = self.w[p + 'conv_filter/kernel'], self.w[p + 'conv_gate/kernel']
            f = old @ wf[0] + now @ wf[1] + self._b(p + 'conv_filter/bias', self.D)
            g = old @ wg[0] + now @ wg[1] + self._b(p + 'conv_gate/bias', self.D)
            if self.G:
                f = f + gvec @ self.w[p + 'gc_filter/kernel'][0]
                g = g + gvec @ self.w[p + 'gc_gate/kernel'][0]
            if self.C:
                lc0 = self.lc_queue[:, 0]                                      # model.py:79-80 keeps index 0
                f = f + lc0 @ self.w[p + 'lc_filter/kernel'][0]
                g = g + lc0 @ self.w[p + 'lc_gate/kernel'][0]
            z = (np.tanh(f) * _sigmoid(g)).astype(F32)                         # model.py:86
            skip = z @ self.w[p + 'skip/kernel'][0] + self._b(p + 'skip/bias', self.S)
            total = skip if total is None else total + skip                    # model.py:157
            cur = (now + (z @ self.w[p + 'dense/kernel'][0] + self._b(p + 'dense/bias', self.R))).astype(F32)
        t1 = np.maximum(total, 0)
        c1 = t1 @ self.w['wavenet/conv1d_1/kernel'][0] + self._b('wavenet/conv1d_1/bias', self.S)
        t2 = np.maximum(c1, 0)
        od = self.O if self.scalar_input else self.Q
        c2 = t2 @ self.w['wavenet/conv1d_2/kernel'][0] + self._b('wavenet/conv1d_2/bias', od)
        return c2.astype(F32)


def sample_mol(y, u):
    """mixture.py:84-114.  y (N,30), u (N,11) uniforms in (1e-5, 1-1e-5)."""
    nr = y.shape[1] // 3
    logit = y[:, :nr]
    sel = np.argmax(logit - np.log(-np.log(u[:, :nr])), axis=1)
    rows = np.arange(y.shape[0])
    means = y[rows, nr + sel]
    ls = np.maximum(y[rows, 2 * nr + sel], F32(np.log(1e-14)))
    u2 = u[:, nr]
    x = means + np.exp(ls) * (np.log(u2) - np.log(F32(1) - u2))
    return np.minimum(np.maximum(x, F32(-1)), F32(1)).astype(F32)


def softmax_f64_to_f32(c2):
    # model.py:243
    z = c2.astype(np.float64)
    z = z - z.max(axis=-1, keepdims=True)
    e = np.exp(z)
    return (e / e.sum(axis=-1, keepdims=True)).astype(F32)


def categorical_draw(prediction, temperature, u):
    """generate.py:219-231 with np.random.choice replaced by its definition
    (cdf = cumsum(float64 p); cdf /= cdf[-1]; searchsorted(cdf, u, 'right'))."""
    with np.errstate(divide='ignore'):
        s = np.log(prediction) / F32(temperature)
    s = s - np.logaddexp.reduce(s, axis=-1, keepdims=True)
    q = np.exp(s)
    out = []
    for p, uu in zip(q, u):
        cdf = np.cumsum(p.astype(np.float64))
        cdf /= cdf[-1]
        out.append(int(np.searchsorted(cdf, uu, side='right')))
    return np.asarray(out), q


def generate(net, T, forced, uniforms, lc_up=None, lc_shift=0, gc_ids=None, temperature=1.0,
             want_logits=False):
    """The per-sample loop of generate.py:202-233 (same contract as OracleModel.generate)."""
    N = net.N
    net.reset_queues()
    forced = np.asarray(forced, F32).reshape(N, -1)
    out = np.zeros((N, T), F32)
    logits = []
    prev = np.zeros(N, F32)
    zero_lc = np.zeros((N, net.C or 1), F32)
    for t in range(T):
        x_in = forced[:, t] if t < forced.shape[1] else prev
        row = None
        if net.C:
            idx = t - lc_shift
            row = lc_up[:, idx] if (lc_up is not None and 0 <= idx < lc_up.shape[1]) else zero_lc
        c2 = net.step(x_in, row, gc_ids)
        if want_logits:
            logits.append(c2)
        if net.scalar_input:
            prev = sample_mol(c2, uniforms[:, t])
        else:
            pred = softmax_f64_to_f32(c2)
            prev, _ = categorical_draw(pred, temperature, uniforms[:, t])
            prev = prev.astype(F32)
        out[:, t] = prev
    if want_logits:
        return out, np.stack(logits, axis=1)
    return out
