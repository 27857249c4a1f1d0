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


def discretized_mix_logistic_loss(*_a, **_k):
    raise NotImplementedError("the MoL loss belongs to the training path (SURVEY.md section 8f, next-3)")
