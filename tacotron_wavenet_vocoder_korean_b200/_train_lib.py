"""ctypes binding of libwn_train_b200.so (include/wn_train_b200.h).  No fallback: if the CUDA library is missing
or no sm_100 device is usable, every entry point raises."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libwn_train_b200.so")

WNT_MAX_LAYERS = 64
WNT_MAX_UPSAMPLE = 8
DTYPES = {'bf16': 0, 'fp32': 1}
WHICH = {'params': 0, 'grads': 1, 'ema': 2, 'adam_m': 3, 'adam_v': 4}


class WntConfig(C.Structure):
    _fields_ = [("batch_size", C.c_int32), ("n_layers", C.c_int32), ("dilations", C.c_int32 * WNT_MAX_LAYERS),
                ("residual_channels", C.c_int32), ("dilation_channels", C.c_int32), ("skip_channels", C.c_int32),
                ("out_channels", C.c_int32), ("quantization_channels", C.c_int32),
                ("use_biases", C.c_int32), ("scalar_input", C.c_int32), ("initial_filter_width", C.c_int32),
                ("gc_channels", C.c_int32), ("gc_cardinality", C.c_int32), ("lc_channels", C.c_int32),
                ("n_upsample", C.c_int32), ("upsample_factor", C.c_int32 * WNT_MAX_UPSAMPLE),
                ("sample_size", C.c_int32), ("dtype", C.c_int32)]


class WntInfo(C.Structure):
    _fields_ = [("n_params", C.c_int64), ("n_weights", C.c_int64), ("n_trainable", C.c_int64), ("workspace_bytes", C.c_int64),
                ("receptive_field", C.c_int32), ("output_width", C.c_int32), ("rows_per_crop", C.c_int32), ("mel_frames", C.c_int32),
                ("gemm_launches", C.c_int64), ("kernel_launches", C.c_int64), ("flops_per_step", C.c_double),
                ("fused_launches", C.c_int64)]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


class WntAdam(C.Structure):
    _fields_ = [("learning_rate", C.c_float), ("beta1", C.c_float), ("beta2", C.c_float), ("epsilon", C.c_float),
                ("ema_decay", C.c_float), ("grad_scale", C.c_float), ("clip_norm", C.c_float), ("t", C.c_int32)]


EXPORTS = ["wnt_create", "wnt_destroy", "wnt_last_error", "wnt_get_info", "wnt_bind", "wnt_set_tensor", "wnt_get_tensor",
           "wnt_variable_names", "wnt_params_changed", "wnt_loss_and_grads", "wnt_apply", "wnt_debug_get"]

_lib = None


def make_config(sample_size, dtype, batch_size, dilations, filter_width, residual_channels, dilation_channels, skip_channels,
                quantization_channels=2 ** 8, out_channels=30, use_biases=False, scalar_input=False, initial_filter_width=32,
                global_condition_channels=None, global_condition_cardinality=None, local_condition_channels=80,
                upsample_factor=None, train_mode=True):
    """WaveNetModel(...) constructor arguments (wavenet/model.py:8-10) -> wnt_config."""
    if filter_width != 2:
        raise NotImplementedError("filter_width != 2 (the reference only uses 2)")
    if len(dilations) > WNT_MAX_LAYERS:
        raise ValueError("at most %d layers" % WNT_MAX_LAYERS)
    if dtype not in DTYPES:
        raise ValueError("dtype must be one of %s" % sorted(DTYPES))
    c = WntConfig()
    c.batch_size, c.n_layers = int(batch_size), len(dilations)
    for i, d in enumerate(dilations):
        c.dilations[i] = int(d)
    c.residual_channels, c.dilation_channels, c.skip_channels = residual_channels, dilation_channels, skip_channels
    c.out_channels, c.quantization_channels = out_channels, quantization_channels
    c.use_biases, c.scalar_input, c.initial_filter_width = int(bool(use_biases)), int(bool(scalar_input)), initial_filter_width
    c.gc_channels = int(global_condition_channels or 0)
    c.gc_cardinality = int(global_condition_cardinality or 0)
    c.lc_channels = int(local_condition_channels or 0)
    up = list(upsample_factor or [])
    if len(up) > WNT_MAX_UPSAMPLE:
        raise ValueError("at most %d upsample stages" % WNT_MAX_UPSAMPLE)
    c.n_upsample = len(up) if c.lc_channels else 0
    for i, f in enumerate(up[:c.n_upsample]):
        c.upsample_factor[i] = int(f)
    c.sample_size = int(sample_size)
    c.dtype = DTYPES[dtype]
    return c


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError("%s is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                               "(there is no CPU fallback)" % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        H, P = C.c_void_p, C.c_void_p
        L.wnt_create.argtypes = [C.POINTER(WntConfig), C.POINTER(H)]
        L.wnt_destroy.argtypes = [H]
        L.wnt_destroy.restype = None
        L.wnt_last_error.argtypes = [H]
        L.wnt_last_error.restype = C.c_char_p
        L.wnt_get_info.argtypes = [H, C.POINTER(WntInfo)]
        L.wnt_bind.argtypes = [H, P, P, P, P, P]
        L.wnt_set_tensor.argtypes = [H, C.c_int, C.c_char_p, P, C.c_int64]
        L.wnt_get_tensor.argtypes = [H, C.c_int, C.c_char_p, P, C.c_int64]
        L.wnt_get_tensor.restype = C.c_int64
        L.wnt_variable_names.argtypes = [H, C.c_char_p, C.c_int64]
        L.wnt_variable_names.restype = C.c_int64
        L.wnt_params_changed.argtypes = [H, P]
        L.wnt_loss_and_grads.argtypes = [H, P, P, P, C.c_float, P, P]
        L.wnt_apply.argtypes = [H, C.POINTER(WntAdam), P]
        L.wnt_debug_get.argtypes = [H, C.c_char_p, P, C.c_int64]
        L.wnt_debug_get.restype = C.c_int64
        _lib = L
    return _lib
