"""Audio helpers of the reference on and next to the WaveNet path: utils/audio.py:14-17 save_wav (host) and
utils/audio.py:69-75 melspectrogram (sm_100a kernel through the C ABI)."""
import numpy as np
from scipy.io import wavfile


def save_wav(wav, path, sr):
    """Peak-normalise to int16 and write (utils/audio.py:14-17).  Unlike the reference this does not
    scale the caller's array in place."""
    wav = np.array(wav, copy=True)                     # same dtype as the caller's array: the reference scales a float32 waveform
    if not np.issubdtype(wav.dtype, np.floating):      # in float32, and the int16 truncation depends on that (tests/test_reference_pin.py)
        wav = wav.astype(np.float32)
    wav *= 32767 / max(0.01, np.max(np.abs(wav)) if wav.size else 0.0)
    wavfile.write(path, sr, wav.astype(np.int16))


def melspectrogram(wav, hparams):
    """utils/audio.py:69-75: wav (n,) or (rows, n) -> (num_mels, frames) [or (rows, num_mels, frames)] float32
    CUDA tensor, the layout the reference returns.  Runs wn_melspectrogram (csrc/wn_mel.cuh); no CPU fallback."""
    import ctypes as C
    import torch
    from . import _lib
    if getattr(hparams, 'use_lws', False):
        raise NotImplementedError("use_lws=True (hparams.py:15) is not supported")
    if not getattr(hparams, 'signal_normalization', True) or not getattr(hparams, 'allow_clipping_in_normalization', True) \
            or not getattr(hparams, 'symmetric_mels', True):
        raise NotImplementedError("only the reference defaults signal_normalization / allow_clipping / symmetric_mels = True")
    if not torch.cuda.is_available():
        raise RuntimeError("melspectrogram: no CUDA device; this package has no CPU fallback")
    w = torch.as_tensor(wav, dtype=torch.float32).cuda().contiguous()
    single = w.dim() == 1
    if single:
        w = w[None]
    rows, n = w.shape
    mc = _lib.WnMelConfig(hparams.sample_rate, hparams.fft_size, hparams.hop_size, hparams.win_size, hparams.num_mels,
                          int(bool(hparams.preemphasize)), hparams.preemphasis, hparams.min_level_db, hparams.ref_level_db,
                          hparams.max_abs_value)
    frames = 1 + n // hparams.hop_size
    out = torch.empty((rows, frames, hparams.num_mels), dtype=torch.float32, device=w.device)
    rc = _lib.lib().wn_melspectrogram(C.c_void_p(w.data_ptr()), rows, n, C.byref(mc), C.c_void_p(out.data_ptr()),
                                      C.c_void_p(torch.cuda.current_stream().cuda_stream))
    if rc != 0:
        raise RuntimeError("wn_melspectrogram: %s (code %d)" % (_lib.lib().wn_last_error(None).decode(), rc))
    out = out.transpose(1, 2)
    return out[0] if single else out
