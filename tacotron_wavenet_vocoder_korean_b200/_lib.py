"""ctypes binding of libwn_b200.so (include/wn_b200.h).  There is no fallback: if the CUDA
library is missing or no sm_100 device is usable, every compute entry point raises."""
import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("WN_B200_LIB") or os.path.join(_HERE, "libwn_b200.so")      # WN_B200_LIB: A/B runs of two builds on one box

WN_MAX_LAYERS = 256
WN_MAX_UPSAMPLE = 8
WN_MAX_BATCH = 32
WN_FLAG_GENERIC_KERNEL = 1
WN_FLAG_NO_DIE_AWARE = 2
WN_FLAG_NO_CLUSTER = 4
WN_FLAG_FAST_ACT = 8


class WnConfig(C.Structure):
    _fields_ = [
        ("batch", C.c_int32), ("n_layers", C.c_int32), ("filter_width", C.c_int32),
        ("residual_channels", C.c_int32), ("dilation_channels", C.c_int32), ("skip_channels", C.c_int32),
        ("quantization_channels", C.c_int32), ("out_channels", C.c_int32),
        ("use_biases", C.c_int32), ("scalar_input", C.c_int32), ("initial_filter_width", C.c_int32),
        ("gc_channels", C.c_int32), ("gc_cardinality", C.c_int32), ("lc_channels", C.c_int32),
        ("n_upsample", C.c_int32), ("upsample_factor", C.c_int32 * WN_MAX_UPSAMPLE),
        ("dilations", C.c_int32 * WN_MAX_LAYERS),
        ("force_M", C.c_int32), ("force_Mt", C.c_int32), ("flags", C.c_int32),
    ]


class WnPlan(C.Structure):
    _fields_ = [(n, C.c_int32) for n in
                ("M", "Mt", "t_cur", "t_old", "t_lc", "t_gc", "t_dense", "t_skip", "t_post1", "t_post2", "t_causal")]

    def as_dict(self):
        return {n: int(getattr(self, n)) for n, _ in self._fields_}


class WnInfo(C.Structure):
    _fields_ = [("grid", C.c_int32), ("threads", C.c_int32), ("M", C.c_int32), ("Mt", C.c_int32),
                ("smem_bytes_layer", C.c_int32), ("smem_bytes_tail", C.c_int32), ("smem_bytes_sampler", C.c_int32),
                ("sm_count", C.c_int32), ("p_hot", C.c_int64), ("weights_in_smem", C.c_int64),
                ("weights_in_global", C.c_int64), ("kernel_launches", C.c_int64),
                ("static_shape", C.c_int32), ("die_aware", C.c_int32), ("cluster_path", C.c_int32), ("fast_act", C.c_int32)]

    def as_dict(self):
        return {n: int(getattr(self, n)) for n, _ in self._fields_}


class WnGenerateArgs(C.Structure):
    _fields_ = [("rows", C.c_int32), ("T", C.c_int32), ("T_row", C.POINTER(C.c_int32)),
                ("n_forced", C.c_int32), ("forced_dev", C.c_void_p), ("lc_dev", C.c_void_p),
                ("t_lc", C.c_int32), ("lc_shift", C.c_int32), ("gc_ids", C.POINTER(C.c_int32)),
                ("uniforms_dev", C.c_void_p), ("temperature", C.c_float),
                ("out_samples_dev", C.c_void_p), ("out_logits_dev", C.c_void_p),
                ("mel_dev", C.c_void_p), ("t_mel", C.c_int32)]


class WnStepArgs(C.Structure):
    _fields_ = [("rows", C.c_int32), ("x_in_dev", C.c_void_p), ("lc_row_dev", C.c_void_p), ("gc_ids", C.POINTER(C.c_int32)),
                ("uniforms_dev", C.c_void_p), ("temperature", C.c_float), ("out_logits_dev", C.c_void_p),
                ("out_probs_dev", C.c_void_p), ("out_sample_dev", C.c_void_p)]


class WnMelConfig(C.Structure):
    _fields_ = [("sample_rate", C.c_int32), ("fft_size", C.c_int32), ("hop_size", C.c_int32), ("win_size", C.c_int32),
                ("num_mels", C.c_int32), ("preemphasize", C.c_int32), ("preemphasis", C.c_float),
                ("min_level_db", C.c_float), ("ref_level_db", C.c_float), ("max_abs_value", C.c_float)]


EXPORTS = ["wn_create", "wn_destroy", "wn_last_error", "wn_set_weight", "wn_finalize", "wn_plan_config", "wn_get_plan",
           "wn_get_info", "wn_receptive_field", "wn_upsample", "wn_generate", "wn_sync_check",
           "wn_generate_host", "wn_mu_law_encode", "wn_mu_law_decode", "wn_melspectrogram",
           "wn_state_create", "wn_state_reset", "wn_state_destroy", "wn_step", "wn_mol_sample", "wn_mol_loss"]

_lib = None


def build(verbose=False):
    """Compile libwn_b200.so in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    cmd = ["make", "-C", os.path.join(_HERE, "csrc")]
    out = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if verbose or out.returncode != 0:
        print(out.stdout)
    if out.returncode != 0:
        raise RuntimeError("building libwn_b200.so failed")
    return LIB_PATH


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError("%s is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                               "(there is no CPU fallback)" % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        H = C.c_void_p
        L.wn_create.argtypes = [C.POINTER(WnConfig), C.POINTER(H)]
        L.wn_destroy.argtypes = [H]
        L.wn_destroy.restype = None
        L.wn_last_error.argtypes = [H]
        L.wn_last_error.restype = C.c_char_p
        L.wn_set_weight.argtypes = [H, C.c_char_p, C.c_void_p, C.c_int64]
        L.wn_finalize.argtypes = [H]
        L.wn_plan_config.argtypes = [C.POINTER(WnConfig), C.c_int, C.POINTER(WnPlan), C.POINTER(WnInfo)]
        L.wn_get_plan.argtypes = [H, C.POINTER(WnPlan)]
        L.wn_get_info.argtypes = [H, C.POINTER(WnInfo)]
        L.wn_receptive_field.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int]
        L.wn_upsample.argtypes = [H, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        L.wn_generate.argtypes = [H, C.POINTER(WnGenerateArgs), C.c_void_p]
        L.wn_sync_check.argtypes = [H, C.c_void_p]
        L.wn_state_create.argtypes = [H, C.c_int, C.POINTER(H)]
        L.wn_state_reset.argtypes = [H, H, C.c_void_p]
        L.wn_state_destroy.argtypes = [H]
        L.wn_state_destroy.restype = None
        L.wn_step.argtypes = [H, H, C.POINTER(WnStepArgs), C.c_void_p]
        L.wn_generate_host.argtypes = [H, C.POINTER(WnGenerateArgs), C.c_void_p, C.c_int]
        L.wn_mu_law_encode.argtypes = [C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_void_p]
        L.wn_mu_law_decode.argtypes = [C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        L.wn_melspectrogram.argtypes = [C.c_void_p, C.c_int, C.c_int64, C.POINTER(WnMelConfig), C.c_void_p, C.c_void_p]
        L.wn_mol_sample.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_float, C.c_void_p, C.c_void_p]
        L.wn_mol_loss.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p]
        _lib = L
    return _lib
