"""wavenet/mixture.py of the reference on CUDA tensors through the C ABI (wn_mol_sample / wn_mol_loss, include/wn_b200.h).
The generation kernels draw in-kernel and the training step evaluates the loss fused with its gradient; this module is the
tensor-level entry point of the reference module for callers that hold logits.  No CPU fallback."""
import ctypes as C
import math

import torch

from .. import _lib
from .ops import _require_cuda, _stream


def sample_from_discretized_mix_logistic(y, log_scale_min=None, uniforms=None):
    """mixture.py:84-114.  y: (B, T, C) conv2 outputs, C = 3 * nr_mix -> (B, T) samples in [-1, 1].  `uniforms`
    (B, T, nr_mix + 1) in (1e-5, 1 - 1e-5) replaces TF's unseeded tf.random_uniform (nr_mix Gumbel draws + one logistic
    draw per step); drawn with torch's generator when absent."""
    if log_scale_min is None:
        log_scale_min = float(math.log(1e-14))
    y = _require_cuda(y, "sample_from_discretized_mix_logistic").to(torch.float32).contiguous()
    assert y.dim() == 3 and y.shape[2] % 3 == 0
    nr = y.shape[2] // 3
    if uniforms is None:
        uniforms = torch.empty(y.shape[0], y.shape[1], nr + 1, device=y.device).uniform_(1e-5, 1.0 - 1e-5)
    u = _require_cuda(uniforms, "sample_from_discretized_mix_logistic").to(torch.float32).contiguous()
    assert tuple(u.shape) == (y.shape[0], y.shape[1], nr + 1), "uniforms must be (B, T, nr_mix + 1)"
    out = torch.empty(y.shape[:2], dtype=torch.float32, device=y.device)
    rc = _lib.lib().wn_mol_sample(C.c_void_p(y.data_ptr()), C.c_void_p(u.data_ptr()), y.shape[0] * y.shape[1], nr,
                                  float(log_scale_min), C.c_void_p(out.data_ptr()), _stream())
    if rc != 0:
        raise RuntimeError("wn_mol_sample failed (%d)" % rc)
    return out


def discretized_mix_logistic_loss(y_hat, y, num_class=256, log_scale_min=None, reduce=True):
    """mixture.py:27-81 for callers that hold network outputs: y_hat (B, T, 3*nr_mix), y (B, T, 1) in [-1, 1] -> the summed loss
    (`reduce=True`) or the per-step losses (B, T), evaluated in fp32 like the reference's graph.  The training step does NOT go
    through this function: libwn_train_b200 evaluates the same formula fused with its gradient (csrc/wn_train_kernels.cuh,
    mol_loss_kernel)."""
    if log_scale_min is None:
        log_scale_min = float(math.log(1e-14))
    y_hat = _require_cuda(y_hat, "discretized_mix_logistic_loss").to(torch.float32).contiguous()
    y = _require_cuda(y, "discretized_mix_logistic_loss").to(torch.float32).contiguous()
    assert y_hat.dim() == 3 and y_hat.shape[2] % 3 == 0
    rows = y_hat.shape[0] * y_hat.shape[1]
    assert y.numel() == rows, "y must be (B, T, 1)"
    nr = y_hat.shape[2] // 3
    per_step = None if reduce else torch.empty(y_hat.shape[:2], dtype=torch.float32, device=y_hat.device)
    total = torch.empty(1, dtype=torch.float64, device=y_hat.device) if reduce else None
    rc = _lib.lib().wn_mol_loss(C.c_void_p(y_hat.data_ptr()), C.c_void_p(y.data_ptr()), rows, nr, int(num_class), float(log_scale_min),
                                C.c_void_p(per_step.data_ptr()) if per_step is not None else None,
                                C.c_void_p(total.data_ptr()) if total is not None else None, _stream())
    if rc != 0:
        raise RuntimeError("wn_mol_loss failed (%d)" % rc)
    return total[0].to(torch.float32) if reduce else per_step
