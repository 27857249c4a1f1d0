"""mu-law codec of the reference's wavenet/ops.py:22-47, on CUDA tensors through the C ABI."""
import ctypes as C

import torch

from .. import _lib


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _require_cuda(t, name):
    if not isinstance(t, torch.Tensor):
        t = torch.as_tensor(t)
    if not t.is_cuda:
        if not torch.cuda.is_available():
            raise RuntimeError("%s: no CUDA device; this package has no CPU fallback" % name)
        t = t.cuda()
    return t


def mu_law_encode(audio, quantization_channels):
    """Quantizes waveform amplitudes (wavenet/ops.py:22-33).  Returns int32 ids."""
    a = _require_cuda(audio, "mu_law_encode").to(torch.float32).contiguous()
    out = torch.empty(a.shape, dtype=torch.int32, device=a.device)
    rc = _lib.lib().wn_mu_law_encode(C.c_void_p(a.data_ptr()), a.numel(), int(quantization_channels),
                                     C.c_void_p(out.data_ptr()), _stream())
    if rc != 0:
        raise RuntimeError("wn_mu_law_encode failed (%d)" % rc)
    return out


def mu_law_decode(output, quantization_channels, quantization=True):
    """Recovers waveform from quantized values (wavenet/ops.py:36-47)."""
    a = _require_cuda(output, "mu_law_decode").to(torch.float32).contiguous()
    out = torch.empty_like(a)
    rc = _lib.lib().wn_mu_law_decode(C.c_void_p(a.data_ptr()), a.numel(), int(quantization_channels),
                                     int(bool(quantization)), C.c_void_p(out.data_ptr()), _stream())
    if rc != 0:
        raise RuntimeError("wn_mu_law_decode failed (%d)" % rc)
    return out


def _adam(learning_rate=1e-3, **_k):
    """wavenet/ops.py:3-5 create_adam_optimizer.  The reference's own trainer never goes through this table (add_optimizer builds
    tf.train.AdamOptimizer directly, wavenet/model.py:325); here Adam is fused into libwn_train_b200's apply step, so the entry only
    carries the hyper-parameters WaveNetTrainer.apply takes."""
    return dict(optimizer='adam', learning_rate=learning_rate, beta1=0.9, beta2=0.999, epsilon=1e-8)


def _unsupported(*_a, **_k):
    raise NotImplementedError("only Adam (the optimizer wavenet/model.py:325 hard-codes) is built into the B200 training step")


# wavenet/ops.py:16-19 exports this mapping.
optimizer_factory = {'adam': _adam, 'sgd': _unsupported, 'rmsprop': _unsupported}
