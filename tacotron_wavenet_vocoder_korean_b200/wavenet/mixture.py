"""wavenet/mixture.py of the reference: the sampler lives inside the persistent kernel
(csrc/wn_kernel.cu, sampler_role); this module exposes it for a single batch of logits."""
import torch


def sample_from_discretized_mix_logistic(y, log_scale_min=None, uniforms=None):
    """y: (B, T, C) conv2 outputs.  Torch restatement of mixture.py:84-114 for callers that hold
    logits (the fused kernel draws in-kernel and never materialises y unless asked).  `uniforms`
    (B, T, C//3 + 1) in (1e-5, 1-1e-5) replaces TF's unseeded RNG."""
    import math
    if log_scale_min is None:
        log_scale_min = float(math.log(1e-14))
    assert y.dim() == 3 and y.shape[2] % 3 == 0
    nr = y.shape[2] // 3
    if uniforms is None:
        uniforms = torch.empty(y.shape[0], y.shape[1], nr + 1, device=y.device).uniform_(1e-5, 1.0 - 1e-5)
    u1, u2 = uniforms[..., :nr], uniforms[..., nr]
    sel = torch.argmax(y[..., :nr] - torch.log(-torch.log(u1)), dim=2, keepdim=True)
    means = torch.gather(y[..., nr:2 * nr], 2, sel).squeeze(2)
    log_scales = torch.clamp(torch.gather(y[..., 2 * nr:3 * nr], 2, sel).squeeze(2), min=log_scale_min)
    x = means + torch.exp(log_scales) * (torch.log(u2) - torch.log(1.0 - u2))
    return torch.clamp(x, -1.0, 1.0)


def discretized_mix_logistic_loss(y_hat, y, num_class=256, log_scale_min=None, reduce=True):
    """mixture.py:27-81 for callers that hold network outputs: y_hat (B, T, 3*nr_mix), y (B, T, 1) in [-1, 1] -> the summed loss
    (`reduce=True`) or the per-step losses (B, T).  The training step does NOT go through this function: libwn_train_b200
    evaluates the same formula fused with its gradient (csrc/wn_train_kernels.cuh, mol_loss_kernel); this is the tensor-level
    entry point of the reference module, evaluated with torch ops on the tensors' device."""
    import math
    import torch.nn.functional as F
    if log_scale_min is None:
        log_scale_min = float(math.log(1e-14))
    assert y_hat.dim() == 3 and y_hat.shape[2] % 3 == 0
    nr = y_hat.shape[2] // 3
    logit_probs, means = y_hat[..., :nr], y_hat[..., nr:2 * nr]
    log_scales = torch.clamp(y_hat[..., 2 * nr:3 * nr], min=log_scale_min)
    y = y.expand(-1, -1, nr)
    centered = y - means
    inv_stdv = torch.exp(-log_scales)
    plus_in = inv_stdv * (centered + 1. / (num_class - 1))
    min_in = inv_stdv * (centered - 1. / (num_class - 1))
    cdf_delta = torch.sigmoid(plus_in) - torch.sigmoid(min_in)
    mid_in = inv_stdv * centered
    log_pdf_mid = mid_in - log_scales - 2. * F.softplus(mid_in)
    inner = torch.where(cdf_delta > 1e-5, torch.log(torch.clamp(cdf_delta, min=1e-12)), log_pdf_mid - math.log((num_class - 1) / 2))
    log_probs = torch.where(y < -0.999, plus_in - F.softplus(plus_in), torch.where(y > 0.999, -F.softplus(min_in), inner))
    log_probs = log_probs + F.log_softmax(logit_probs, -1)
    lse = torch.logsumexp(log_probs, -1)
    return -lse.sum() if reduce else -lse
