"""ctypes binding of libtaco_b200.so (include/taco_b200.h).  No fallback: if the CUDA library is missing or
no sm_100 device is usable, every compute entry point raises."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("TACO_B200_LIB") or os.path.join(_HERE, "libtaco_b200.so")     # TACO_B200_LIB: A/B runs of two builds on one box

TACO_MAX_PRENET = 4
TACO_MAX_PROJ = 4
TACO_MAX_DEC_LAYERS = 4
ATTENTION_TYPES = {'bah_mon': 0, 'bah_mon_norm': 1, 'loc_sen': 2}


class TacoConfig(C.Structure):
    _fields_ = [
        ("num_symbols", C.c_int32), ("embedding_size", C.c_int32), ("num_speakers", C.c_int32),
        ("speaker_embedding_size", C.c_int32),
        ("n_enc_prenet", C.c_int32), ("enc_prenet_sizes", C.c_int32 * TACO_MAX_PRENET),
        ("enc_bank_size", C.c_int32), ("enc_bank_channel_size", C.c_int32), ("enc_highway_depth", C.c_int32),
        ("enc_rnn_size", C.c_int32),
        ("n_enc_proj", C.c_int32), ("enc_proj_sizes", C.c_int32 * TACO_MAX_PROJ), ("enc_proj_width", C.c_int32),
        ("attention_type", C.c_int32), ("attention_size", C.c_int32), ("attention_state_size", C.c_int32),
        ("dec_layer_num", C.c_int32), ("dec_rnn_size", C.c_int32),
        ("n_dec_prenet", C.c_int32), ("dec_prenet_sizes", C.c_int32 * TACO_MAX_PRENET),
        ("post_bank_size", C.c_int32), ("post_bank_channel_size", C.c_int32), ("post_highway_depth", C.c_int32),
        ("post_rnn_size", C.c_int32),
        ("n_post_proj", C.c_int32), ("post_proj_sizes", C.c_int32 * TACO_MAX_PROJ), ("post_proj_width", C.c_int32),
        ("reduction_factor", C.c_int32), ("max_iters", C.c_int32), ("num_mels", C.c_int32), ("num_freq", C.c_int32),
    ]


class TacoInfo(C.Structure):
    _fields_ = [("sm_count", C.c_int32), ("dec_grid", C.c_int32), ("dec_threads", C.c_int32),
                ("dec_smem_bytes", C.c_int32), ("dec_phases_per_step", C.c_int32), ("rnn_weights_in_smem", C.c_int32),
                ("n_params", C.c_int64), ("kernel_launches", C.c_int64), ("workspace_bytes", C.c_int64),
                ("tc_gemm_launches", C.c_int64)]

    def as_dict(self):
        return {n: int(getattr(self, n)) for n, _ in self._fields_}


class TacoSynthArgs(C.Structure):
    _fields_ = [("N", C.c_int32), ("T_in", C.c_int32), ("ids_dev", C.c_void_p), ("lengths", C.POINTER(C.c_int32)),
                ("speaker_ids", C.POINTER(C.c_int32)), ("n_steps", C.c_int32), ("manual_alignments_dev", C.c_void_p),
                ("mel_dev", C.c_void_p), ("linear_dev", C.c_void_p), ("alignments_dev", C.c_void_p)]


EXPORTS = ["taco_create", "taco_destroy", "taco_last_error", "taco_set_weight", "taco_finalize", "taco_get_info",
           "taco_synthesize", "taco_sync_check", "taco_synthesize_host", "taco_debug_get"]

_lib = None


def make_config(hp, num_speakers):
    """hp: mapping or attribute bag with the Tacotron fields of hparams.py:124-158."""
    get = (lambda k: hp[k]) if isinstance(hp, dict) else (lambda k: getattr(hp, k))
    cfg = TacoConfig()

    def put_list(name, count_field, limit):
        vals = list(get(name))
        if len(vals) > limit:
            raise ValueError("%s: at most %d entries" % (name, limit))
        setattr(cfg, count_field, len(vals))
        arr = getattr(cfg, name)
        for i, v in enumerate(vals):
            arr[i] = int(v)
    at = get('attention_type')
    if at not in ATTENTION_TYPES:
        raise NotImplementedError("attention_type %r: only %s are built" % (at, sorted(ATTENTION_TYPES)))
    if num_speakers > 1 and get('model_type') != 'deepvoice':
        raise NotImplementedError("multi-speaker model_type %r: only 'deepvoice' (hparams.py:137) is built" % get('model_type'))
    try:
        cfg.num_symbols = int(get('num_symbols'))
    except (KeyError, AttributeError):
        cfg.num_symbols = 80
    cfg.embedding_size = get('embedding_size')
    cfg.num_speakers = num_speakers
    cfg.speaker_embedding_size = get('speaker_embedding_size')
    put_list('enc_prenet_sizes', 'n_enc_prenet', TACO_MAX_PRENET)
    cfg.enc_bank_size = get('enc_bank_size')
    cfg.enc_bank_channel_size = get('enc_bank_channel_size')
    if get('enc_maxpool_width') != 2 or get('post_maxpool_width') != 2:
        raise NotImplementedError("maxpool_width != 2 (hparams.py:148,166)")
    cfg.enc_highway_depth = get('enc_highway_depth')
    cfg.enc_rnn_size = get('enc_rnn_size')
    put_list('enc_proj_sizes', 'n_enc_proj', TACO_MAX_PROJ)
    cfg.enc_proj_width = get('enc_proj_width')
    cfg.attention_type = ATTENTION_TYPES[at]
    cfg.attention_size = get('attention_size')
    cfg.attention_state_size = get('attention_state_size')
    cfg.dec_layer_num = get('dec_layer_num')
    cfg.dec_rnn_size = get('dec_rnn_size')
    put_list('dec_prenet_sizes', 'n_dec_prenet', TACO_MAX_PRENET)
    cfg.post_bank_size = get('post_bank_size')
    cfg.post_bank_channel_size = get('post_bank_channel_size')
    cfg.post_highway_depth = get('post_highway_depth')
    cfg.post_rnn_size = get('post_rnn_size')
    put_list('post_proj_sizes', 'n_post_proj', TACO_MAX_PROJ)
    cfg.post_proj_width = get('post_proj_width')
    cfg.reduction_factor = get('reduction_factor')
    cfg.max_iters = get('max_iters')
    cfg.num_mels = get('num_mels')
    cfg.num_freq = get('num_freq')
    return cfg


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError("%s is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                               "(there is no CPU fallback)" % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        H = C.c_void_p
        L.taco_create.argtypes = [C.POINTER(TacoConfig), C.POINTER(H)]
        L.taco_destroy.argtypes = [H]
        L.taco_destroy.restype = None
        L.taco_last_error.argtypes = [H]
        L.taco_last_error.restype = C.c_char_p
        L.taco_set_weight.argtypes = [H, C.c_char_p, C.c_void_p, C.c_int64]
        L.taco_finalize.argtypes = [H]
        L.taco_get_info.argtypes = [H, C.POINTER(TacoInfo)]
        L.taco_synthesize.argtypes = [H, C.POINTER(TacoSynthArgs), C.c_void_p]
        L.taco_sync_check.argtypes = [H, C.c_void_p]
        L.taco_synthesize_host.argtypes = [H, C.POINTER(TacoSynthArgs)]
        L.taco_debug_get.argtypes = [H, C.c_char_p, C.c_void_p, C.c_int64]
        L.taco_debug_get.restype = C.c_int64
        _lib = L
    return _lib
