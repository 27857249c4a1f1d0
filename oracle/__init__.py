"""TEST INFRASTRUCTURE -- ctypes front-end of the CPU oracle (oracle/wn_oracle.c).

PARITY: pinned to the reference's own Python run on a numpy TF stand-in, TF kernels unpinned -- see the header of wn_oracle.c.  Only tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference legs may import this package; the
product package never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

ORC_MAX_LAYERS = 256
ORC_MAX_UP = 8


class OrcConfig(C.Structure):
    _fields_ = [
        ("batch", C.c_int32), ("n_layers", C.c_int32), ("filter_width", C.c_int32),
        ("residual_channels", C.c_int32), ("dilation_channels", C.c_int32), ("skip_channels", C.c_int32),
        ("quantization_channels", C.c_int32), ("out_channels", C.c_int32),
        ("use_biases", C.c_int32), ("scalar_input", C.c_int32), ("initial_filter_width", C.c_int32),
        ("gc_channels", C.c_int32), ("gc_cardinality", C.c_int32), ("lc_channels", C.c_int32),
        ("n_upsample", C.c_int32), ("upsample_factor", C.c_int32 * ORC_MAX_UP),
        ("dilations", C.c_int32 * ORC_MAX_LAYERS),
    ]


class OrcPlan(C.Structure):
    _fields_ = [(n, C.c_int32) for n in
                ("M", "Mt", "t_cur", "t_old", "t_lc", "t_gc", "t_dense", "t_skip", "t_post1", "t_post2", "t_causal")]

    @classmethod
    def natural(cls):
        return cls(*([1] * 11))

    @classmethod
    def from_dict(cls, d):
        return cls(*[int(d[n]) for n, _ in cls._fields_])


def build(force=False):
    """Compile liborc.so in place (gcc, a second or two)."""
    so = os.path.join(_HERE, "liborc.so")
    srcs = [os.path.join(_HERE, f) for f in ("wn_oracle.c", "wn_math_ref.h", "wn_cpu_best.h")]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        L.orc_create.restype = C.c_void_p
        L.orc_create.argtypes = [C.POINTER(OrcConfig)]
        L.orc_destroy.argtypes = [C.c_void_p]
        L.orc_last_error.restype = C.c_char_p
        L.orc_last_error.argtypes = [C.c_void_p]
        L.orc_set_weight.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p, C.c_long]
        L.orc_upsample.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        L.orc_generate.argtypes = [C.c_void_p, C.POINTER(OrcPlan), C.c_int, C.c_int, C.c_void_p,
                                   C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_float,
                                   C.c_void_p, C.c_void_p]
        L.orc_generate_best.argtypes = L.orc_generate.argtypes + [C.c_int]
        L.orc_gate_probe.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.orc_gate_probe.restype = None
        L.orc_receptive_field.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int]
        L.orc_mu_law_encode.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        L.orc_mu_law_decode.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]
        L.orc_softmax_probs.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.orc_mol_sample.argtypes = [C.c_void_p, C.c_void_p, C.c_long, C.c_int, C.c_void_p]
        L.orc_mol_sample.restype = None
        L.orc_math_probe.restype = C.c_float
        L.orc_math_probe.argtypes = [C.c_int, C.c_float]
        _LIB = L
    return _LIB


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def make_config(batch_size, dilations, filter_width, residual_channels, dilation_channels, skip_channels,
                quantization_channels=256, out_channels=30, use_biases=False, scalar_input=False,
                initial_filter_width=32, global_condition_channels=None, global_condition_cardinality=None,
                local_condition_channels=80, upsample_factor=None, **_ignored):
    """Same keyword names as WaveNetModel.__init__ (wavenet/model.py:8-10)."""
    cfg = OrcConfig()
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
    uf = list(upsample_factor or [])
    cfg.n_upsample = len(uf)
    for i, f in enumerate(uf):
        cfg.upsample_factor[i] = f
    for i, d in enumerate(dilations):
        cfg.dilations[i] = d
    return cfg


class OracleModel:
    """CPU oracle with the reference's construction signature."""

    def __init__(self, **kwargs):
        self.cfg = make_config(**kwargs)
        self._h = lib().orc_create(C.byref(self.cfg))
        self.out_dim = self.cfg.out_channels if self.cfg.scalar_input else self.cfg.quantization_channels
        self.receptive_field = receptive_field(self.cfg.filter_width, list(kwargs["dilations"]),
                                               bool(self.cfg.scalar_input), self.cfg.initial_filter_width)

    def __del__(self):
        try:
            if self._h:
                lib().orc_destroy(self._h)
                self._h = None
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            raise RuntimeError(lib().orc_last_error(self._h).decode())

    def set_weights(self, state):
        for name, arr in state.items():
            a = np.ascontiguousarray(arr, dtype=np.float32)
            self._check(lib().orc_set_weight(self._h, name.encode(), _ptr(a), a.size))

    def upsample(self, mel):
        """mel (N, T_mel, C) -> (N, T_mel*prod(factors), C); wavenet/model.py:102-111."""
        mel = np.ascontiguousarray(mel, dtype=np.float32)
        n, tm, c = mel.shape
        f = 1
        for i in range(self.cfg.n_upsample):
            f *= self.cfg.upsample_factor[i]
        out = np.empty((n, tm * f, c), np.float32)
        for b in range(n):
            self._check(lib().orc_upsample(self._h, _ptr(mel[b]), tm, _ptr(out[b])))
        return out

    def generate(self, T, forced, uniforms, lc_up=None, lc_shift=0, gc_ids=None, temperature=1.0,
                 plan=None, want_logits=False, best_effort_threads=0):
        """Run T steps per batch row.  forced: (N, n_forced) fp32, n_forced>=1.
        uniforms: (N,T,nr_mix+1) fp32 for scalar input, (N,T) fp64 for mu-law."""
        n = self.cfg.batch
        forced = np.ascontiguousarray(forced, dtype=np.float32).reshape(n, -1)
        if self.cfg.scalar_input:
            uniforms = np.ascontiguousarray(uniforms, dtype=np.float32)
            assert uniforms.shape == (n, T, self.cfg.out_channels // 3 + 1), uniforms.shape
        else:
            uniforms = np.ascontiguousarray(uniforms, dtype=np.float64)
            assert uniforms.shape == (n, T), uniforms.shape
        t_lc = 0
        if lc_up is not None:
            lc_up = np.ascontiguousarray(lc_up, dtype=np.float32)
            t_lc = lc_up.shape[1]
        gc = np.ascontiguousarray(gc_ids, dtype=np.int32) if gc_ids is not None else None
        out = np.empty((n, T), np.float32)
        logits = np.empty((n, T, self.out_dim), np.float32) if want_logits else None
        plan = plan or OrcPlan.natural()
        if best_effort_threads:
            # wn_cpu_best.h: weight-stationary threads, all rows together; bit-identical to orc_generate for the same plan
            self._check(lib().orc_generate_best(self._h, C.byref(plan), T, forced.shape[1], _ptr(forced), _ptr(lc_up),
                                                t_lc, lc_shift, _ptr(gc), _ptr(uniforms), float(temperature),
                                                _ptr(out), _ptr(logits), int(best_effort_threads)))
        else:
            self._check(lib().orc_generate(self._h, C.byref(plan), T, forced.shape[1], _ptr(forced), _ptr(lc_up),
                                           t_lc, lc_shift, _ptr(gc), _ptr(uniforms), float(temperature),
                                           _ptr(out), _ptr(logits)))
        return (out, logits) if want_logits else out


def receptive_field(filter_width, dilations, scalar_input, initial_filter_width):
    d = np.asarray(dilations, dtype=np.int32)
    return lib().orc_receptive_field(filter_width, _ptr(d), len(d), int(scalar_input), initial_filter_width)


def mu_law_encode(audio, quantization_channels):
    a = np.ascontiguousarray(audio, dtype=np.float32)
    out = np.empty(a.shape, np.int32)
    lib().orc_mu_law_encode(_ptr(a), a.size, quantization_channels, _ptr(out))
    return out


def mu_law_decode(output, quantization_channels, quantization=True):
    a = np.ascontiguousarray(output, dtype=np.float32)
    out = np.empty(a.shape, np.float32)
    lib().orc_mu_law_decode(_ptr(a), a.size, quantization_channels, int(quantization), _ptr(out))
    return out


def mol_sample(y, uniforms):
    """wavenet/mixture.py:84-114 on (…, 3*nr_mix) logits with (…, nr_mix + 1) uniforms -> (…) samples."""
    y = np.ascontiguousarray(y, dtype=np.float32)
    u = np.ascontiguousarray(uniforms, dtype=np.float32)
    nr = y.shape[-1] // 3
    assert y.shape[-1] == 3 * nr and u.shape == y.shape[:-1] + (nr + 1,)
    out = np.empty(y.shape[:-1], np.float32)
    lib().orc_mol_sample(_ptr(y), _ptr(u), out.size, nr, _ptr(out))
    return out


def gate_probe(f, g):
    """z = tanh32(f) * sigmoid32(g) by the 8-lane path of wn_cpu_best.h and by the scalar pinned functions."""
    f = np.ascontiguousarray(f, dtype=np.float32)
    g = np.ascontiguousarray(g, dtype=np.float32)
    zv, zs = np.empty(f.shape, np.float32), np.empty(f.shape, np.float32)
    lib().orc_gate_probe(_ptr(f), _ptr(g), f.size, _ptr(zv), _ptr(zs))
    return zv, zs


def softmax_probs(logits_row):
    a = np.ascontiguousarray(logits_row, dtype=np.float32)
    out = np.empty_like(a)
    lib().orc_softmax_probs(_ptr(a), a.size, _ptr(out))
    return out


def math_probe(which, x):
    return float(lib().orc_math_probe(which, float(x)))
