# coding: utf-8
"""WaveNet vocoder training step on B200 (BASELINE configs[3]).

`WaveNetTrainer` is the eager counterpart of the graph the reference builds in train_vocoder.py:100-123:
`net.add_loss(...)` (wavenet/model.py:247-312) + `net.add_optimizer(hparams, global_step)` (:314-346) and the loop body
`sess.run([global_step, loss, optimize])` (train_vocoder.py:169).  Compute is libwn_train_b200.so (cuBLASLt bf16 GEMMs +
hand-written kernels, include/wn_train_b200.h); torch holds the flat parameter / gradient / Adam / EMA buffers so that
data-parallel training is ONE `all_reduce` of one tensor per step (NCCL over NVLink; the reference has no data
parallelism at all).  No CPU fallback.
"""
import ctypes as C

import numpy as np
import torch

from .. import _train_lib


def learning_rate_at(hparams, global_step):
    """tf.train.exponential_decay(wavenet_learning_rate, global_step, wavenet_decay_steps, wavenet_decay_rate)
    (wavenet/model.py:321), non-staircase."""
    get = (lambda k: hparams[k]) if isinstance(hparams, dict) else (lambda k: getattr(hparams, k))
    return float(get('wavenet_learning_rate')) * float(get('wavenet_decay_rate')) ** (float(global_step) / float(get('wavenet_decay_steps')))


class WaveNetTrainer(object):
    def __init__(self, sample_size, dtype='bf16', device=None, **model_kwargs):
        if not torch.cuda.is_available():
            raise RuntimeError("no CUDA device: the B200 WaveNet training path has no CPU fallback")
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.kw = dict(model_kwargs)
        self.sample_size = int(sample_size)
        self.dtype = dtype
        L = _train_lib.lib()
        self._cfg = _train_lib.make_config(sample_size, dtype, **model_kwargs)
        self._h = C.c_void_p()
        with torch.cuda.device(self.device):
            rc = L.wnt_create(C.byref(self._cfg), C.byref(self._h))
            if rc != 0:
                raise ValueError("wnt_create: %s (code %d)" % (L.wnt_last_error(None).decode(), rc))
            i = self.info()
            n = int(i['n_params'])
            self.params, self.grads, self.adam_m, self.adam_v, self.ema = (torch.zeros(n, dtype=torch.float32, device=self.device) for _ in range(5))
            self._check(L.wnt_bind(self._h, *[C.c_void_p(t.data_ptr()) for t in (self.params, self.grads, self.adam_m, self.adam_v, self.ema)]))
        self._loss = torch.zeros(1, dtype=torch.float32, device=self.device)
        self.global_step = 0
        buf = C.create_string_buffer(int(L.wnt_variable_names(self._h, None, 0)) + 1)
        L.wnt_variable_names(self._h, buf, len(buf))
        self.variable_names = buf.value.decode().split('\n')

    def __del__(self):
        try:
            if getattr(self, "_h", None) and self._h.value:
                torch.cuda.synchronize(self.device)
                _train_lib.lib().wnt_destroy(self._h)
                self._h = C.c_void_p()
        except Exception:
            pass

    def _check(self, rc):
        if rc < 0:
            raise RuntimeError("libwn_train_b200: %s (code %d)" % (_train_lib.lib().wnt_last_error(self._h).decode(), rc))
        return rc

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def info(self):
        i = _train_lib.WntInfo()
        self._check(_train_lib.lib().wnt_get_info(self._h, C.byref(i)))
        return i.as_dict()

    # ---- Saver.restore / Saver.save -------------------------------------------------------------------------------------
    def load_state_dict(self, state, which='params', init_ema=True):
        """{TF variable name: array} -> the flat buffer.  Loading the parameters also initialises the EMA shadows to the
        same values, as tf.train.ExponentialMovingAverage.apply does when it creates them (wavenet/model.py:346)."""
        L = _train_lib.lib()
        w = _train_lib.WHICH[which]
        with torch.cuda.device(self.device):
            torch.cuda.synchronize()
            for name in self.variable_names:
                if name not in state:
                    raise KeyError("variable %r missing from the state dict" % name)
                a = np.ascontiguousarray(np.asarray(state[name], dtype=np.float32))
                self._check(L.wnt_set_tensor(self._h, w, name.encode(), a.ctypes.data_as(C.c_void_p), a.size))
            if which == 'params':
                self._check(L.wnt_params_changed(self._h, self._stream()))
                if init_ema:
                    self.ema.copy_(self.params)
        return self

    def state_dict(self, which='params'):
        L = _train_lib.lib()
        w = _train_lib.WHICH[which]
        shapes = self.variable_shapes()
        out = {}
        with torch.cuda.device(self.device):
            for name in self.variable_names:
                a = np.empty(shapes[name], np.float32)
                self._check(L.wnt_get_tensor(self._h, w, name.encode(), a.ctypes.data_as(C.c_void_p), a.size))
                out[name] = a
        return out

    def variable_shapes(self):
        from .. import synth
        shapes = synth.weight_shapes(**self.kw)
        return {k: tuple(shapes[k]) for k in self.variable_names}

    def _check_gc_range(self, ids):
        """Speaker ids index the embedding table on the device, so they are range-checked first.  Host data is checked on the
        host; a CUDA tensor is checked once per (storage, version) -- reading it back costs a device synchronisation, which
        a training loop that reuses its id tensor must not pay every step."""
        card = self._cfg.gc_cardinality
        if torch.is_tensor(ids) and ids.is_cuda:
            key = (ids.data_ptr(), ids._version, tuple(ids.shape))
            if getattr(self, '_gc_checked', None) == key:
                return
            lo, hi = int(ids.min()), int(ids.max())
            self._gc_checked = key
        else:
            a = np.asarray(ids.cpu() if torch.is_tensor(ids) else ids)
            lo, hi = int(a.min()), int(a.max())
        if lo < 0 or hi >= card:
            self._gc_checked = None
            raise ValueError("global condition id out of range [0, %d)" % card)

    # ---- the step ---------------------------------------------------------------------------------------------------
    def loss_and_grads(self, input_batch, local_condition=None, global_condition_batch=None, l2_regularization_strength=None):
        """add_loss + compute_gradients.  input_batch (N, sample_size[, 1]) float in [-1, 1]; local_condition
        (N, sample_size / hop, num_mels); global_condition_batch (N,) speaker ids.  Returns the loss as a 1-element CUDA
        tensor (no host sync); gradients land in self.grads."""
        N = self._cfg.batch_size
        with torch.cuda.device(self.device):
            wav = torch.as_tensor(input_batch, dtype=torch.float32, device=self.device).reshape(N, -1).contiguous()
            if wav.shape[1] != self.sample_size:
                raise ValueError("input_batch has %d samples per crop, the trainer was built for %d" % (wav.shape[1], self.sample_size))
            mel = gc = None
            if self._cfg.lc_channels:
                if local_condition is None:
                    raise ValueError("local_condition is required (local_condition_channels=%d)" % self._cfg.lc_channels)
                mel = torch.as_tensor(local_condition, dtype=torch.float32, device=self.device).contiguous()
                frames = self.info()['mel_frames']
                if tuple(mel.shape) != (N, frames, self._cfg.lc_channels):
                    raise ValueError("local_condition must be %s, got %s" % ((N, frames, self._cfg.lc_channels), tuple(mel.shape)))
            if self._cfg.gc_channels:
                if global_condition_batch is None:
                    raise ValueError("global_condition_batch is required (global_condition_channels=%d)" % self._cfg.gc_channels)
                self._check_gc_range(global_condition_batch)
                gc = torch.as_tensor(global_condition_batch, device=self.device).reshape(N).to(torch.int32).contiguous()
            l2 = -1.0 if l2_regularization_strength is None else float(l2_regularization_strength)
            self._keep = (wav, mel, gc)
            self._check(_train_lib.lib().wnt_loss_and_grads(
                self._h, C.c_void_p(wav.data_ptr()), C.c_void_p(mel.data_ptr()) if mel is not None else None,
                C.c_void_p(gc.data_ptr()) if gc is not None else None, C.c_float(l2), C.c_void_p(self._loss.data_ptr()), self._stream()))
        return self._loss

    def apply(self, learning_rate, t=None, grad_scale=1.0, clip_norm=0.0, beta1=0.9, beta2=0.999, epsilon=1e-8, ema_decay=0.9999):
        """optimizer.apply_gradients + ema.apply (wavenet/model.py:333-346); increments global_step."""
        # the bias-correction step is Adam's own counter (beta1_power / beta2_power in the TF checkpoint), which the reference's
        # Saver restores independently of a global_step reset (train_vocoder.py --restore_from with a new logdir)
        if getattr(self, 'adam_t', None) is None:
            self.adam_t = self.global_step
        a = _train_lib.WntAdam(learning_rate, beta1, beta2, epsilon, ema_decay, grad_scale, clip_norm, int(t if t is not None else self.adam_t + 1))
        with torch.cuda.device(self.device):
            self._check(_train_lib.lib().wnt_apply(self._h, C.byref(a), self._stream()))
        self.global_step += 1
        self.adam_t += 1

    def train_step(self, input_batch, local_condition, global_condition_batch, hparams, l2_regularization_strength=None):
        """One `sess.run([global_step, loss, optimize])` (train_vocoder.py:169).  Under torch.distributed the gradients
        (and the reported loss) are averaged over ranks with one all_reduce of the flat buffer."""
        from .. import dist as wdist
        get = (lambda k, d=None: hparams.get(k, d)) if isinstance(hparams, dict) else (lambda k, d=None: getattr(hparams, k, d))
        loss = self.loss_and_grads(input_batch, local_condition, global_condition_batch, l2_regularization_strength)
        scale, loss = wdist.allreduce_mean_(self.grads, loss)
        lr = learning_rate_at(hparams, self.global_step)
        self.apply(lr, grad_scale=scale, clip_norm=1.0 if get('wavenet_clip_gradients', False) else 0.0)
        return loss

    def sync_params(self, src=0):
        """Broadcast rank `src`'s parameters / optimizer state to every rank and refresh the compute copy."""
        from .. import dist as wdist
        for t in (self.params, self.adam_m, self.adam_v, self.ema):
            wdist.broadcast_flat_(t, src)
        with torch.cuda.device(self.device):
            self._check(_train_lib.lib().wnt_params_changed(self._h, self._stream()))

    def debug_get(self, name):
        L = _train_lib.lib()
        n = self._check(L.wnt_debug_get(self._h, name.encode(), None, 0))
        out = np.empty(int(n), np.float32)
        self._check(L.wnt_debug_get(self._h, name.encode(), out.ctypes.data_as(C.c_void_p), out.size))
        return out
