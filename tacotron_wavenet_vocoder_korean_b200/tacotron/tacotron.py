"""Tacotron with the reference's construction / initialize signatures (tacotron/tacotron.py:31-37) whose
inference graph runs on sm_100a kernels through libtaco_b200.so (include/taco_b200.h).

In the reference `initialize` builds a symbolic TF graph over placeholders and `sess.run` evaluates
`linear_outputs, alignments, mel_outputs` (synthesizer.py:129-160).  Here `initialize` receives the actual
batch (token ids, lengths, speaker ids) and evaluates eagerly: afterwards `.mel_outputs (N, T, num_mels)`,
`.linear_outputs (N, T, num_freq)` and `.alignments (N, T_in, T_dec)` are torch CUDA tensors.  Weights are
exchanged as a dict keyed by TF variable names (synth.taco_weight_shapes).  Training entry points
(`add_loss`, `add_optimizer`, tacotron.py:258-300) are out of scope (inference only).
There is no CPU fallback.
"""
import ctypes as C

import numpy as np
import torch

from .. import _taco_lib


class Tacotron(object):
    def __init__(self, hparams):
        self._hparams = hparams
        self._h = None
        self._weights = None
        self._num_speakers = None
        self.is_manual_attention = False      # tacotron.py:122-123 placeholders -> plain attributes
        self.manual_alignments = None

    # ---- weights -------------------------------------------------------------------------------------
    def load_state_dict(self, weights):
        """weights: {tf_variable_name: array}.  (Saver.restore in the reference, synthesizer.py:68-70.)"""
        self._weights = {k: np.ascontiguousarray(np.asarray(v, dtype=np.float32)) for k, v in weights.items()}
        self._destroy()

    def _destroy(self):
        if self._h is not None:
            _taco_lib.lib().taco_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self._destroy()
        except Exception:
            pass

    def _err(self, what, rc):
        L = _taco_lib.lib()
        msg = L.taco_last_error(self._h).decode() if self._h is not None else L.taco_last_error(None).decode()
        return RuntimeError("%s: %s (code %d)" % (what, msg, rc))

    def _ensure_handle(self, num_speakers):
        if self._h is not None and self._num_speakers == num_speakers:
            return
        self._destroy()
        if self._weights is None:
            raise RuntimeError("Tacotron: load_state_dict() must be called before initialize()")
        if not torch.cuda.is_available():
            raise RuntimeError("Tacotron: no CUDA device; this package has no CPU fallback")
        L = _taco_lib.lib()
        cfg = _taco_lib.make_config(self._hparams, num_speakers)
        h = C.c_void_p()
        rc = L.taco_create(C.byref(cfg), C.byref(h))
        if rc != 0:
            raise ValueError("taco_create: %s" % L.taco_last_error(None).decode())
        self._h = h
        self._num_speakers = num_speakers
        for name, arr in self._weights.items():
            rc = L.taco_set_weight(h, name.encode(), arr.ctypes.data_as(C.c_void_p), arr.size)
            if rc != 0:
                raise self._err("taco_set_weight(%s)" % name, rc)
        rc = L.taco_finalize(h)
        if rc != 0:
            raise self._err("taco_finalize", rc)

    def info(self):
        inf = _taco_lib.TacoInfo()
        _taco_lib.lib().taco_get_info(self._h, C.byref(inf))
        return inf.as_dict()

    # ---- the graph -----------------------------------------------------------------------------------
    def initialize(self, inputs, input_lengths, num_speakers, speaker_id, mel_targets=None, linear_targets=None,
                   loss_coeff=None, rnn_decoder_test_mode=False, is_randomly_initialized=False, n_steps=None,
                   want_linear=True):
        """tacotron/tacotron.py:36-235 with rnn_decoder_test_mode=True (synthesizer.py:56).
        inputs (N, T_in) int token ids, input_lengths (N), speaker_id (N) or None."""
        if mel_targets is not None or linear_targets is not None or not rnn_decoder_test_mode:
            raise NotImplementedError("training graph (teacher forcing, dropout, batch-norm updates) is out of scope; "
                                      "call initialize(..., rnn_decoder_test_mode=True)")
        hp = self._hparams
        get = (lambda k: hp[k]) if isinstance(hp, dict) else (lambda k: getattr(hp, k))
        self._ensure_handle(int(num_speakers))
        dev = torch.device('cuda', torch.cuda.current_device())
        ids = torch.as_tensor(np.asarray(inputs), dtype=torch.int32).to(dev).contiguous()
        N, T_in = ids.shape
        lens = np.ascontiguousarray(np.asarray(input_lengths, dtype=np.int32).reshape(N))
        spk = None if speaker_id is None else np.ascontiguousarray(np.asarray(speaker_id, dtype=np.int32).reshape(N))
        S = int(n_steps or get('max_iters'))
        r, nm, nf = get('reduction_factor'), get('num_mels'), get('num_freq')
        mel = torch.empty((N, S * r, nm), dtype=torch.float32, device=dev)
        lin = torch.empty((N, S * r, nf), dtype=torch.float32, device=dev) if want_linear else None
        al = torch.empty((N, T_in, S), dtype=torch.float32, device=dev)
        man = None
        if self.is_manual_attention:
            man = torch.as_tensor(np.asarray(self.manual_alignments), dtype=torch.float32).to(dev).contiguous()
            if tuple(man.shape) != (N, S, T_in):
                raise ValueError("manual_alignments must be (N, n_steps, T_in) = %r, got %r" % ((N, S, T_in), tuple(man.shape)))
        a = _taco_lib.TacoSynthArgs()
        a.N, a.T_in, a.n_steps = N, T_in, S
        a.ids_dev = ids.data_ptr()
        a.lengths = lens.ctypes.data_as(C.POINTER(C.c_int32))
        a.speaker_ids = spk.ctypes.data_as(C.POINTER(C.c_int32)) if spk is not None else None
        a.manual_alignments_dev = man.data_ptr() if man is not None else None
        a.mel_dev = mel.data_ptr()
        a.linear_dev = lin.data_ptr() if lin is not None else None
        a.alignments_dev = al.data_ptr()
        rc = _taco_lib.lib().taco_synthesize(self._h, C.byref(a), C.c_void_p(torch.cuda.current_stream().cuda_stream))
        if rc != 0:
            raise self._err("taco_synthesize", rc)
        rc = _taco_lib.lib().taco_sync_check(self._h, C.c_void_p(torch.cuda.current_stream().cuda_stream))
        if rc != 0:
            raise self._err("taco_sync_check", rc)
        # TacoTestHelper (helpers.py:35-41): stop after the first step at which every sentence has emitted an
        # all-zero output; dynamic_decode keeps evaluating finished rows, so truncating afterwards is equivalent.
        if n_steps is None:
            fin = (mel.view(N, S, r * nm) == 0).all(dim=2).to(torch.int32).cummax(dim=1).values.bool().all(dim=0)
            if bool(fin.any()):
                stop = int(torch.nonzero(fin)[0]) + 1
                if stop < S:
                    return self.initialize(inputs, input_lengths, num_speakers, speaker_id, rnn_decoder_test_mode=True,
                                           n_steps=stop, want_linear=want_linear)
        self.inputs = ids
        self.speaker_id = speaker_id
        self.input_lengths = lens
        self.loss_coeff = loss_coeff
        self.mel_outputs = mel
        self.linear_outputs = lin
        self.alignments = al
        self.mel_targets = None
        self.linear_targets = None
        self.num_speakers = num_speakers
        return self

    def debug_tensor(self, name, shape):
        """Test hook over taco_debug_get: an intermediate of the last initialize() as a numpy array."""
        if any(d < 0 for d in shape):       # one free dimension: ask the library for the size first
            total = _taco_lib.lib().taco_debug_get(self._h, name.encode(), None, 0)
            if total < 0:
                raise RuntimeError("no debug tensor %s" % name)
            known = int(np.prod([d for d in shape if d >= 0]))
            shape = tuple(d if d >= 0 else total // max(known, 1) for d in shape)
        out = np.empty(shape, np.float32)
        n = _taco_lib.lib().taco_debug_get(self._h, name.encode(), out.ctypes.data_as(C.c_void_p), out.size)
        if n != out.size:
            raise RuntimeError("debug tensor %s has %d floats, asked for %d" % (name, n, out.size))
        return out

    def add_loss(self):
        raise NotImplementedError("Tacotron training (tacotron.py:258-283) is out of scope (SURVEY.md section 8)")

    def add_optimizer(self, global_step):
        raise NotImplementedError("Tacotron training (tacotron.py:285-320) is out of scope (SURVEY.md section 8)")

    def get_dummy_feed_dict(self):
        return {'is_manual_attention': False, 'manual_alignments': np.zeros([1, 1, 1])}
