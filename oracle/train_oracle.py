# coding: utf-8
"""TEST INFRASTRUCTURE -- torch (CPU, fp32/fp64, autograd) restatement of the reference's WaveNet TRAINING graph
(SURVEY.md section 8f next-3).  Only tests/, __graft_entry__.smoke() and the CPU legs of the benchmarks may import it.

PARITY: the loss value is pinned to the reference's OWN add_loss (wavenet/model.py:247-312, run unmodified on the numpy TF
stand-in tests/golden/tf_numpy_shim.py -> tests/golden/ref_train.npz, reproduced to 2e-5 incl. the L2 term and the per-layer
lc alignment); TensorFlow's kernels and its autodiff are not installable here and stay unpinned.  Also pinned in
tests/test_train_oracle.py: the MoL loss against an independent float64 numpy evaluation of mixture.py:27-81, the
network forward against the incremental-generation oracle (teacher forcing: training logits at position t == the
per-sample oracle's logits), Adam/decay/EMA against torch.optim.Adam and the closed forms.

Follows (reference file:line), in the reference's own left-aligned slicing so that it is structurally independent of
the CUDA implementation (which indexes every layer by absolute time):
  wavenet/model.py:247-312   add_loss: scalar input or mu-law one-hot, cut last sample (:267-271), create_upsample (:276),
                             _create_network (:278), target = input[rf:] (:286), MoL loss mean (:289-290) or softmax
                             cross-entropy (:293-296), optional L2 on non-bias variables (:303-308)
  wavenet/model.py:112-167   _create_network (train mode): causal conv 'valid', dilation stack, sum of skips, post 1x1s
  wavenet/model.py:66-101    _create_dilation_layer: valid dilated conv, gc/lc added (lc sliced from index 0, :79-80),
                             tanh*sigmoid, dense residual on the input cut from the left (:98-101), skip on the last
                             output_width steps (:94-96)
  wavenet/model.py:102-111   create_upsample
  wavenet/mixture.py:27-81   discretized_mix_logistic_loss(num_class=2**16, reduce=False)
  wavenet/model.py:314-346   add_optimizer: exponential_decay, Adam, optional clip_by_global_norm(1.), EMA(0.9999)
"""
import numpy as np
import torch
import torch.nn.functional as F

LOG_SCALE_MIN = float(np.log(1e-14))


def mol_loss(y_hat, y, num_class=2 ** 16, log_scale_min=LOG_SCALE_MIN):
    """mixture.py:27-81 with reduce=False: y_hat (B,T,3K), y (B,T,1) -> (B,T)."""
    K = y_hat.shape[2] // 3
    logit_probs = y_hat[:, :, :K]
    means = y_hat[:, :, K:2 * K]
    log_scales = torch.clamp(y_hat[:, :, 2 * K:3 * K], min=log_scale_min)
    y = y.expand(-1, -1, K)
    centered = y - means
    inv_stdv = torch.exp(-log_scales)
    plus_in = inv_stdv * (centered + 1. / (num_class - 1))
    cdf_plus = torch.sigmoid(plus_in)
    min_in = inv_stdv * (centered - 1. / (num_class - 1))
    cdf_min = torch.sigmoid(min_in)
    log_cdf_plus = plus_in - F.softplus(plus_in)
    log_one_minus_cdf_min = -F.softplus(min_in)
    cdf_delta = cdf_plus - cdf_min
    mid_in = inv_stdv * centered
    log_pdf_mid = mid_in - log_scales - 2. * F.softplus(mid_in)
    inner = torch.where(cdf_delta > 1e-5, torch.log(torch.clamp(cdf_delta, min=1e-12)),
                        log_pdf_mid - float(np.log((num_class - 1) / 2)))
    log_probs = torch.where(y < -0.999, log_cdf_plus, torch.where(y > 0.999, log_one_minus_cdf_min, inner))
    log_probs = log_probs + F.log_softmax(logit_probs, -1)
    return -torch.logsumexp(log_probs, -1)


def mol_loss_np(y_hat, y, num_class=2 ** 16, log_scale_min=LOG_SCALE_MIN):
    """Independent float64 numpy evaluation of the same formula (pins mol_loss)."""
    y_hat = np.asarray(y_hat, np.float64)
    y = np.asarray(y, np.float64)
    K = y_hat.shape[-1] // 3
    lp, mu, ls = y_hat[..., :K], y_hat[..., K:2 * K], np.maximum(y_hat[..., 2 * K:], log_scale_min)
    yk = np.repeat(y, K, axis=-1)
    c = yk - mu
    inv = np.exp(-ls)
    h = 1.0 / (num_class - 1)

    def softplus(x):
        return np.logaddexp(0.0, x)

    def sig(x):
        return 1.0 / (1.0 + np.exp(-x))
    p, m, mid = inv * (c + h), inv * (c - h), inv * c
    delta = sig(p) - sig(m)
    out = np.where(yk < -0.999, p - softplus(p),
                   np.where(yk > 0.999, -softplus(m),
                            np.where(delta > 1e-5, np.log(np.maximum(delta, 1e-12)),
                                     mid - ls - 2 * softplus(mid) - np.log((num_class - 1) / 2))))
    lsm = lp - lp.max(-1, keepdims=True)
    lsm = lsm - np.log(np.exp(lsm).sum(-1, keepdims=True))
    a = out + lsm
    mx = a.max(-1)
    return -(mx + np.log(np.exp(a - mx[..., None]).sum(-1)))


def mu_law_encode_t(x, Q):
    """wavenet/ops.py:22-33."""
    mu = float(Q - 1)
    safe = torch.clamp(torch.abs(x), max=1.0)
    mag = torch.log1p(mu * safe) / float(np.log1p(mu))
    return ((torch.sign(x) * mag + 1) / 2 * mu + 0.5).to(torch.int64)


class TorchWaveNetTrain(object):
    def __init__(self, weights, batch_size, dilations, filter_width, residual_channels, dilation_channels, skip_channels,
                 quantization_channels=256, out_channels=30, use_biases=False, scalar_input=False, initial_filter_width=32,
                 global_condition_channels=None, global_condition_cardinality=None, local_condition_channels=80,
                 upsample_factor=None, train_mode=True, dtype=torch.float32):
        assert filter_width == 2
        self.N, self.dil = batch_size, list(dilations)
        self.R, self.D, self.S = residual_channels, dilation_channels, skip_channels
        self.Q, self.O = quantization_channels, out_channels
        self.use_biases, self.scalar, self.ifw = use_biases, scalar_input, initial_filter_width
        self.G, self.card, self.C = global_condition_channels, global_condition_cardinality, local_condition_channels
        self.up = list(upsample_factor or [])
        self.rf = sum(self.dil) + 1 + ((self.ifw - 1) if self.scalar else 1)
        self.dtype = dtype
        self.p = {k: torch.tensor(np.asarray(v), dtype=dtype, requires_grad=True) for k, v in weights.items()}

    def _conv(self, x, w, dilation=1):
        # x (N,T,Cin), TF kernel (k, Cin, Cout) -> 'valid' conv1d
        return F.conv1d(x.transpose(1, 2), w.permute(2, 1, 0), dilation=dilation).transpose(1, 2)

    def create_upsample(self, mel):
        x = mel[:, None]                                                    # N,1,T,C
        for i, f in enumerate(self.up):
            k = self.p['wavenet/upsample%d/kernel' % i].reshape(1, 1, f, 2)
            x = F.conv_transpose2d(x, k, stride=(f, 1))[..., :mel.shape[2]]  # pinned in tests/test_oracle.py
        return x[:, 0]

    def network(self, inp, lc, gvec):
        p = self.p
        cur = self._conv(inp, p['wavenet/conv1d/kernel'])
        out_w = inp.shape[1] - self.rf + 1
        total = None
        for l, d in enumerate(self.dil):
            pre = 'wavenet/dilated_stack/layer%d/dilation_layer/' % l
            f = self._conv(cur, p[pre + 'conv_filter/kernel'], d)
            g = self._conv(cur, p[pre + 'conv_gate/kernel'], d)
            if self.use_biases:
                f = f + p[pre + 'conv_filter/bias']
                g = g + p[pre + 'conv_gate/bias']
            if gvec is not None:
                f = f + gvec @ p[pre + 'gc_filter/kernel'][0]
                g = g + gvec @ p[pre + 'gc_gate/kernel'][0]
            if lc is not None:
                f = f + (lc @ p[pre + 'lc_filter/kernel'][0])[:, :f.shape[1]]
                g = g + (lc @ p[pre + 'lc_gate/kernel'][0])[:, :g.shape[1]]
            z = torch.tanh(f) * torch.sigmoid(g)
            tr = z @ p[pre + 'dense/kernel'][0]
            sk = z[:, z.shape[1] - out_w:] @ p[pre + 'skip/kernel'][0]
            if self.use_biases:
                tr = tr + p[pre + 'dense/bias']
                sk = sk + p[pre + 'skip/bias']
            cur = cur[:, cur.shape[1] - tr.shape[1]:] + tr
            total = sk if total is None else total + sk
        h = torch.relu(total) @ p['wavenet/conv1d_1/kernel'][0]
        if self.use_biases:
            h = h + p['wavenet/conv1d_1/bias']
        o = torch.relu(h) @ p['wavenet/conv1d_2/kernel'][0]
        if self.use_biases:
            o = o + p['wavenet/conv1d_2/bias']
        return o

    def raw_output(self, wav, mel=None, gc_ids=None):
        wav = torch.as_tensor(np.asarray(wav), dtype=self.dtype)
        if self.scalar:
            net_in = wav[:, :, None]
        else:
            ids = mu_law_encode_t(wav, self.Q)
            net_in = F.one_hot(ids, self.Q).to(self.dtype)
        inp = net_in[:, :-1]
        lc = self.create_upsample(torch.as_tensor(np.asarray(mel), dtype=self.dtype)) if mel is not None else None
        gvec = self.p['wavenet/gc_embedding'][torch.as_tensor(np.asarray(gc_ids), dtype=torch.int64)][:, None] if self.G else None
        return self.network(inp, lc, gvec), net_in

    def loss(self, wav, mel=None, gc_ids=None, l2_regularization_strength=None):
        raw, net_in = self.raw_output(wav, mel, gc_ids)
        target = net_in[:, self.rf:]
        if self.scalar:
            red = mol_loss(raw, target).mean()
        else:
            logp = F.log_softmax(raw.reshape(-1, self.Q), -1)
            red = -(target.reshape(-1, self.Q) * logp).sum(-1).mean()
        if l2_regularization_strength is None:
            return red
        l2 = sum((v ** 2).sum() / 2 for k, v in self.p.items() if 'bias' not in k)
        return red + l2_regularization_strength * l2

    def loss_and_grads(self, wav, mel=None, gc_ids=None, l2_regularization_strength=None):
        for v in self.p.values():
            v.grad = None
        L = self.loss(wav, mel, gc_ids, l2_regularization_strength)
        L.backward()
        grads = {k: (v.grad.detach().numpy().copy() if v.grad is not None else np.zeros(tuple(v.shape), np.float32))
                 for k, v in self.p.items()}
        return float(L.detach()), grads


def learning_rate(hp_lr, global_step, decay_steps, decay_rate):
    """tf.train.exponential_decay (non-staircase), model.py:321."""
    return hp_lr * decay_rate ** (global_step / float(decay_steps))


def adam_ema_step(params, grads, m, v, ema, t, lr, clip=False, beta1=0.9, beta2=0.999, eps=1e-8, ema_decay=0.9999):
    """One `sess.run(optimize)` of model.py:314-346 in float64 numpy, in place.  t = number of this update (1-based).
    tf.train.AdamOptimizer: lr_t = lr*sqrt(1-b2^t)/(1-b1^t); p -= lr_t*m/(sqrt(v)+eps).  EMA: s -= (1-decay)*(s-p)."""
    if clip:
        gn = np.sqrt(sum(float((g.astype(np.float64) ** 2).sum()) for g in grads.values()))
        scale = 1.0 / max(gn, 1.0)
        grads = {k: g * scale for k, g in grads.items()}
    lr_t = lr * np.sqrt(1 - beta2 ** t) / (1 - beta1 ** t)
    for k in params:
        g = grads[k].astype(np.float64)
        m[k] = beta1 * m[k] + (1 - beta1) * g
        v[k] = beta2 * v[k] + (1 - beta2) * g * g
        params[k] = params[k] - lr_t * m[k] / (np.sqrt(v[k]) + eps)
        ema[k] = ema[k] - (1 - ema_decay) * (ema[k] - params[k])
