"""WaveNetModel with the reference's construction signature (wavenet/model.py:8-10) whose
generation path runs as one persistent sm_100a kernel through libwn_b200.so.

Tensors are torch CUDA tensors; weights are exchanged as a dict keyed by the reference's TF
variable names.  There is no CPU fallback.
"""
import os
import ctypes as C

import numpy as np
import torch

from .. import _lib


def _make_cfg(batch_size, dilations, filter_width, residual_channels, dilation_channels, skip_channels,
              quantization_channels=2 ** 8, out_channels=30, use_biases=False, scalar_input=False,
              initial_filter_width=32, global_condition_channels=None, global_condition_cardinality=None,
              local_condition_channels=80, upsample_factor=None, train_mode=True, force_M=0, force_Mt=0,
              generic_kernel=False, die_aware=True, cluster=True, fast_act=False, **_ignored):
    dilations = list(dilations)
    uf = list(upsample_factor) if upsample_factor else []
    if len(uf) > _lib.WN_MAX_UPSAMPLE or len(dilations) > _lib.WN_MAX_LAYERS:
        raise ValueError("too many upsample stages / layers")
    cfg = _lib.WnConfig()
    cfg.batch = batch_size
    cfg.n_layers = len(dilations)
    cfg.filter_width = filter_width
    cfg.residual_channels = residual_channels
    cfg.dilation_channels = dilation_channels
    cfg.skip_channels = skip_channels
    cfg.quantization_channels = quantization_channels
    cfg.out_channels = out_channels
    cfg.use_biases = int(bool(use_biases))
    cfg.scalar_input = int(bool(scalar_input))
    cfg.initial_filter_width = initial_filter_width
    cfg.gc_channels = global_condition_channels or 0
    cfg.gc_cardinality = global_condition_cardinality or 0
    cfg.lc_channels = local_condition_channels or 0
    cfg.n_upsample = len(uf)
    for i, f in enumerate(uf):
        cfg.upsample_factor[i] = f
    for i, d in enumerate(dilations):
        cfg.dilations[i] = d
    cfg.force_M = force_M
    cfg.force_Mt = force_Mt
    cfg.flags = ((_lib.WN_FLAG_GENERIC_KERNEL if generic_kernel else 0) | (0 if die_aware else _lib.WN_FLAG_NO_DIE_AWARE) |
                 (0 if cluster else _lib.WN_FLAG_NO_CLUSTER) | (_lib.WN_FLAG_FAST_ACT if fast_act else 0))
    return cfg


def plan_config(sm_count=148, **model_kwargs):
    """(plan, info) the library picks for a WaveNetModel(**model_kwargs) on a device with `sm_count`
    SMs -- host-only (wn_plan_config), usable without a GPU."""
    cfg = _make_cfg(**model_kwargs)
    plan, info = _lib.WnPlan(), _lib.WnInfo()
    rc = _lib.lib().wn_plan_config(C.byref(cfg), sm_count, C.byref(plan), C.byref(info))
    if rc != 0:
        raise ValueError(_lib.lib().wn_last_error(None).decode())
    return plan.as_dict(), info.as_dict()


class WaveNetModel(object):
    def __init__(self, batch_size, dilations, filter_width, residual_channels, dilation_channels, skip_channels,
                 quantization_channels=2 ** 8, out_channels=30, use_biases=False, scalar_input=False,
                 initial_filter_width=32, global_condition_channels=None, global_condition_cardinality=None,
                 local_condition_channels=80, upsample_factor=None, train_mode=True, device=None,
                 force_M=0, force_Mt=0, generic_kernel=False, die_aware=True, cluster=True, fast_act=False):
        self.batch_size = batch_size
        self.dilations = list(dilations)
        self.filter_width = filter_width
        self.residual_channels = residual_channels
        self.dilation_channels = dilation_channels
        self.quantization_channels = quantization_channels
        self.use_biases = use_biases
        self.skip_channels = skip_channels
        self.scalar_input = scalar_input
        self.initial_filter_width = initial_filter_width
        self.global_condition_channels = global_condition_channels
        self.global_condition_cardinality = global_condition_cardinality
        self.local_condition_channels = local_condition_channels
        self.upsample_factor = list(upsample_factor) if upsample_factor else None
        self.train_mode = train_mode
        self.receptive_field = WaveNetModel.calculate_receptive_field(
            self.filter_width, self.dilations, self.scalar_input, self.initial_filter_width)
        self.out_channels = out_channels
        self.out_dim = out_channels if scalar_input else quantization_channels
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device()) \
            if torch.cuda.is_available() else None

        cfg = _make_cfg(batch_size, dilations, filter_width, residual_channels, dilation_channels, skip_channels,
                        quantization_channels, out_channels, use_biases, scalar_input, initial_filter_width,
                        global_condition_channels, global_condition_cardinality, local_condition_channels,
                        upsample_factor, train_mode, force_M, force_Mt, generic_kernel, die_aware, cluster, fast_act)
        self._cfg = cfg
        self._h = C.c_void_p()
        rc = _lib.lib().wn_create(C.byref(cfg), C.byref(self._h))
        if rc != 0:
            raise ValueError(_lib.lib().wn_last_error(None).decode())
        self._finalized = False
        self._inc = None            # wn_state of predict_proba_incremental: the persistent device queues
        # attribute the reference's generate.py runs through sess.run before the loop (generate.py:163);
        # the kernel zeroes its queues at every launch, so this is a no-op kept for drop-in callers.
        self.queue_initializer = lambda: None

    # ------------------------------------------------------------------ plumbing
    def __del__(self):
        try:
            if getattr(self, "_inc", None) is not None and self._inc.value:
                _lib.lib().wn_state_destroy(self._inc)
                self._inc = None
            if getattr(self, "_h", None) and self._h.value:
                _lib.lib().wn_destroy(self._h)
                self._h = C.c_void_p()
        except Exception:
            pass

    def _err(self):
        return _lib.lib().wn_last_error(self._h).decode()

    def _check(self, rc):
        if rc != 0:
            raise RuntimeError("libwn_b200: %s (code %d)" % (self._err(), rc))

    def _dev_guard(self):
        if self.device is None or not torch.cuda.is_available():
            raise RuntimeError("no CUDA device: the B200 WaveNet path has no CPU fallback")
        return torch.cuda.device(self.device)

    @staticmethod
    def calculate_receptive_field(filter_width, dilations, scalar_input, initial_filter_width):
        """wavenet/model.py:31-39."""
        d = np.ascontiguousarray(dilations, dtype=np.int32)
        return int(_lib.lib().wn_receptive_field(filter_width, d.ctypes.data_as(C.c_void_p), len(d),
                                                 int(bool(scalar_input)), initial_filter_width))

    def load_state_dict(self, state):
        """Restore of the non-queue variables (generate.py:157-161): {tf_variable_name: array}."""
        if self.train_mode:
            self._state = {k: np.array(v.detach().cpu().numpy() if isinstance(v, torch.Tensor) else v, dtype=np.float32)
                           for k, v in state.items() if 'queue' not in k and 'ExponentialMovingAverage' not in k}
            self._trainer = None
        with self._dev_guard():
            for name, arr in state.items():
                if 'queue' in name or 'ExponentialMovingAverage' in name or name.startswith('optimizer'):
                    continue       # generate.py:157 restores only the raw, non-queue variables
                a = np.ascontiguousarray(arr.detach().cpu().numpy() if isinstance(arr, torch.Tensor) else arr,
                                         dtype=np.float32)
                self._check(_lib.lib().wn_set_weight(self._h, name.encode(), a.ctypes.data_as(C.c_void_p), a.size))
            self._check(_lib.lib().wn_finalize(self._h))
        self._finalized = True
        return self

    def plan(self):
        p = _lib.WnPlan()
        self._check(_lib.lib().wn_get_plan(self._h, C.byref(p)))
        return p.as_dict()

    def info(self):
        i = _lib.WnInfo()
        self._check(_lib.lib().wn_get_info(self._h, C.byref(i)))
        return i.as_dict()

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    # ------------------------------------------------------------------ reference methods
    def create_upsample(self, local_condition_batch):
        """wavenet/model.py:102-111: (N, T_mel, C) -> (N, T_mel*prod(upsample_factor), C)."""
        with self._dev_guard():
            mel = torch.as_tensor(local_condition_batch, dtype=torch.float32, device=self.device).contiguous()
            n, tm, c = mel.shape
            f = int(np.prod(self.upsample_factor))
            out = torch.empty((n, tm * f, c), dtype=torch.float32, device=self.device)
            self._check(_lib.lib().wn_upsample(self._h, C.c_void_p(mel.data_ptr()), n, tm,
                                               C.c_void_p(out.data_ptr()), self._stream()))
            return out

    def generate(self, T, forced, uniforms, lc_up=None, lc_shift=0, gc_ids=None, temperature=1.0,
                 want_logits=False, T_row=None, sync=True, mel=None):
        """The fused generate.py:202-233 loop.  All tensor arguments live on the GPU (moved if not).

        forced   (rows, n_forced>=1): x_in(t) for t < n_forced; afterwards the drawn sample feeds back.
        uniforms (rows, T, nr_mix+1) fp32 [scalar input] or (rows, T) fp64 [mu-law].
        lc_up    (rows, T_lc, C) upsampled local condition or None; lc_shift as in include/wn_b200.h.
        mel      (rows, T_mel, C) mel frames instead of lc_up: create_upsample is evaluated inside the kernel
                 (frames staged by TMA), nothing is materialised.
        Returns samples (rows, T) fp32 [and logits (rows, T, out_dim)]."""
        if not self._finalized:
            raise RuntimeError("load_state_dict() must be called before generate()")
        with self._dev_guard():
            dev = self.device
            forced = torch.as_tensor(forced, dtype=torch.float32, device=dev)
            rows = forced.shape[0]
            forced = forced.reshape(rows, -1).contiguous()
            if self.scalar_input:
                uniforms = torch.as_tensor(uniforms, dtype=torch.float32, device=dev).contiguous()
                if tuple(uniforms.shape) != (rows, T, self.out_channels // 3 + 1):
                    raise ValueError("uniforms must be (rows, T, %d)" % (self.out_channels // 3 + 1))
            else:
                uniforms = torch.as_tensor(uniforms, dtype=torch.float64, device=dev).contiguous()
                if tuple(uniforms.shape) != (rows, T):
                    raise ValueError("uniforms must be (rows, T) float64")
            a = _lib.WnGenerateArgs()
            a.rows, a.T, a.n_forced = rows, int(T), forced.shape[1]
            a.forced_dev = forced.data_ptr()
            keep = [forced, uniforms]
            if lc_up is not None:
                lc_up = torch.as_tensor(lc_up, dtype=torch.float32, device=dev).contiguous()
                a.lc_dev, a.t_lc = lc_up.data_ptr(), lc_up.shape[1]
                keep.append(lc_up)
            if mel is not None:
                mel = torch.as_tensor(mel, dtype=torch.float32, device=dev).contiguous()
                a.mel_dev, a.t_mel = mel.data_ptr(), mel.shape[1]
                keep.append(mel)
            a.lc_shift = int(lc_shift)
            if gc_ids is not None:
                g = (C.c_int32 * rows)(*[int(v) for v in gc_ids])
                a.gc_ids = C.cast(g, C.POINTER(C.c_int32))
                keep.append(g)
            if T_row is not None:
                tr = (C.c_int32 * rows)(*[int(v) for v in T_row])
                a.T_row = C.cast(tr, C.POINTER(C.c_int32))
                keep.append(tr)
            a.uniforms_dev = uniforms.data_ptr()
            a.temperature = float(temperature)
            out = torch.empty((rows, T), dtype=torch.float32, device=dev)
            a.out_samples_dev = out.data_ptr()
            logits = None
            if want_logits:
                logits = torch.empty((rows, T, self.out_dim), dtype=torch.float32, device=dev)
                a.out_logits_dev = logits.data_ptr()
            self._check(_lib.lib().wn_generate(self._h, C.byref(a), self._stream()))
            if sync:
                self._check(_lib.lib().wn_sync_check(self._h, self._stream()))
            self._keep = keep
            return (out, logits) if want_logits else out

    def sync_check(self):
        self._check(_lib.lib().wn_sync_check(self._h, self._stream()))

    def generate_host(self, T, forced, uniforms, mel=None, lc_shift=0, gc_ids=None, temperature=1.0):
        """Host-buffer entry point (wn_generate_host): numpy in, numpy out, copies inside."""
        if not self._finalized:
            raise RuntimeError("load_state_dict() must be called before generate_host()")
        with self._dev_guard():
            forced = np.ascontiguousarray(forced, dtype=np.float32)
            rows = forced.shape[0]
            forced = forced.reshape(rows, -1)
            uniforms = np.ascontiguousarray(uniforms, dtype=np.float32 if self.scalar_input else np.float64)
            a = _lib.WnGenerateArgs()
            a.rows, a.T, a.n_forced = rows, int(T), forced.shape[1]
            a.forced_dev = forced.ctypes.data
            a.lc_shift = int(lc_shift)
            g = None
            if gc_ids is not None:
                g = (C.c_int32 * rows)(*[int(v) for v in gc_ids])
                a.gc_ids = C.cast(g, C.POINTER(C.c_int32))
            a.uniforms_dev = uniforms.ctypes.data
            a.temperature = float(temperature)
            out = np.empty((rows, T), np.float32)
            a.out_samples_dev = out.ctypes.data
            mel_p, t_mel = None, 0
            if mel is not None:
                mel = np.ascontiguousarray(mel, dtype=np.float32)
                mel_p, t_mel = mel.ctypes.data_as(C.c_void_p), mel.shape[1]
            self._check(_lib.lib().wn_generate_host(self._h, C.byref(a), mel_p, t_mel))
            return out

    def predict_proba_incremental(self, waveform, upsampled_local_condition=None, global_condition=None,
                                  name='wavenet', uniforms=None, temperature=1.0, return_draw=False):
        """The graph node of wavenet/model.py:215-245, evaluated eagerly: feed ONE step, get the softmax probabilities
        (N, Q) or, for scalar input, a sample (N, 1) -- what the reference's loop obtains from sess.run (generate.py:211).

        The causal / local-condition / dilation queues of model.py:49-64 are persistent device buffers (wn_state) that
        every call advances by one position (wn_step): O(1) per call, any number of calls.  `reset_incremental()` is
        sess.run(net.queue_initializer) (generate.py:163).  For scalar input the node draws from the mixture of logistics
        itself (mixture.py:84-114); `uniforms` (N, nr_mix + 1) makes the draw reproducible, otherwise it is random like
        TF's.  `return_draw=True` (one-hot models, with uniforms (N,) float64) also returns the id generate.py:219-231
        would draw on the host."""
        dev = self.device
        lib = _lib.lib()
        with self._dev_guard():
            if self._inc is None:
                h = C.c_void_p()
                self._check(lib.wn_state_create(self._h, self.batch_size, C.byref(h)))
                self._inc = h
            N = self.batch_size
            x = torch.as_tensor(waveform, dtype=torch.float32, device=dev).reshape(N, -1)[:, -1].contiguous()
            a = _lib.WnStepArgs()
            a.rows = N
            a.x_in_dev = x.data_ptr()
            keep = [x]
            if self.local_condition_channels:
                if upsampled_local_condition is None:
                    raise ValueError("upsampled_local_condition is required (model.py:125)")
                lc = torch.as_tensor(upsampled_local_condition, dtype=torch.float32, device=dev).reshape(N, self.local_condition_channels).contiguous()
                a.lc_row_dev = lc.data_ptr()
                keep.append(lc)
            if self.global_condition_channels:
                gids = [int(v) for v in (global_condition if global_condition is not None else [0] * N)]
                g = (C.c_int32 * N)(*gids)
                a.gc_ids = C.cast(g, C.POINTER(C.c_int32))
                keep.append(g)
            a.temperature = float(temperature)
            if self.scalar_input:
                nr1 = self.out_channels // 3 + 1
                u = uniforms if uniforms is not None else torch.empty(N, nr1, device=dev).uniform_(1e-5, 1 - 1e-5)
                u = torch.as_tensor(u, dtype=torch.float32, device=dev).reshape(N, nr1).contiguous()
                out = torch.empty((N, 1), dtype=torch.float32, device=dev)
                a.uniforms_dev, a.out_sample_dev = u.data_ptr(), out.data_ptr()
                keep += [u]
                self._check(lib.wn_step(self._h, self._inc, C.byref(a), self._stream()))
                self._keep_step = keep
                return out
            probs = torch.empty((N, self.quantization_channels), dtype=torch.float32, device=dev)
            a.out_probs_dev = probs.data_ptr()
            draw = None
            if return_draw:
                if uniforms is None:
                    raise ValueError("return_draw needs uniforms (N,) float64")
                u = torch.as_tensor(uniforms, dtype=torch.float64, device=dev).reshape(N).contiguous()
                draw = torch.empty((N,), dtype=torch.float32, device=dev)
                a.uniforms_dev, a.out_sample_dev = u.data_ptr(), draw.data_ptr()
                keep.append(u)
            self._check(lib.wn_step(self._h, self._inc, C.byref(a), self._stream()))
            self._keep_step = keep
            return (probs, draw) if return_draw else probs

    def reset_incremental(self):
        """Counterpart of sess.run(net.queue_initializer), generate.py:163."""
        if self._inc is not None:
            self._check(_lib.lib().wn_state_reset(self._h, self._inc, self._stream()))

    # ------------------------------------------------------------------ training
    def trainer(self, sample_size, dtype='bf16', init_seed=None):
        """The libwn_train_b200 step for crops of `sample_size` samples, created on first use and seeded with the loaded
        state dict (or TF-default initialisers: Glorot-uniform kernels incl. the upsamplers, zero biases --
        tf.global_variables_initializer, train_vocoder.py:129-130; `init_seed` None draws a fresh seed per run as TF does)."""
        from .train import WaveNetTrainer
        from .. import synth
        tr = getattr(self, '_trainer', None)
        if tr is None or tr.sample_size != int(sample_size) or tr.dtype != dtype:
            kw = dict(batch_size=self.batch_size, dilations=self.dilations, filter_width=self.filter_width,
                      residual_channels=self.residual_channels, dilation_channels=self.dilation_channels,
                      skip_channels=self.skip_channels, quantization_channels=self.quantization_channels,
                      out_channels=self.out_channels, use_biases=self.use_biases, scalar_input=self.scalar_input,
                      initial_filter_width=self.initial_filter_width, global_condition_channels=self.global_condition_channels,
                      global_condition_cardinality=self.global_condition_cardinality,
                      local_condition_channels=self.local_condition_channels, upsample_factor=self.upsample_factor)
            state = tr.state_dict() if tr is not None else getattr(self, '_state', None)
            new = WaveNetTrainer(sample_size, dtype=dtype, device=self.device, **kw)
            if state is None:
                seed = int(init_seed) if init_seed is not None else int.from_bytes(os.urandom(4), 'little')
                state = synth.make_weights(seed=seed, init='train', **kw)
            new.load_state_dict(state)
            if tr is not None:
                new.global_step = tr.global_step
            self._trainer = new
        return self._trainer

    def add_loss(self, input_batch, local_condition=None, global_condition_batch=None, l2_regularization_strength=None,
                 name='wavenet', dtype='bf16'):
        """wavenet/model.py:247-312, evaluated eagerly together with its gradients (the reference builds a symbolic loss and
        lets `optimizer.compute_gradients` differentiate it, :327).  Sets and returns `self.loss` (1-element CUDA tensor)."""
        if not self.train_mode:
            raise RuntimeError("add_loss needs WaveNetModel(train_mode=True)")
        x = torch.as_tensor(input_batch)
        tr = self.trainer(x.reshape(self.batch_size, -1).shape[1], dtype)
        self.loss = tr.loss_and_grads(x, local_condition, global_condition_batch, l2_regularization_strength)
        return self.loss

    def add_optimizer(self, hparams, global_step=None):
        """wavenet/model.py:314-346: exponential-decay Adam (+ optional clip_by_global_norm(1.)) followed by the EMA update.
        Sets `self.optimize`, a callable applying ONE update from the gradients of the last add_loss (`sess.run(net.optimize)`,
        train_vocoder.py:169); `global_step` (an int) overrides the trainer's own counter."""
        from .train import learning_rate_at
        if getattr(self, '_trainer', None) is None:
            raise RuntimeError("add_optimizer supposes that add_loss has been called (wavenet/model.py:315)")
        tr = self._trainer
        if global_step is not None:
            tr.global_step = int(global_step)
        get = (lambda k, d=None: hparams.get(k, d)) if isinstance(hparams, dict) else (lambda k, d=None: getattr(hparams, k, d))
        clip = 1.0 if get('wavenet_clip_gradients', False) else 0.0

        def optimize(grad_scale=1.0):
            self.learning_rate = learning_rate_at(hparams, tr.global_step)
            tr.apply(self.learning_rate, grad_scale=grad_scale, clip_norm=clip)
            return tr.global_step
        self.optimize = optimize
        return optimize
