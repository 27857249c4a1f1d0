"""TEST INFRASTRUCTURE -- numpy restatement of the reference's STFT -> mel chain
(utils/audio.py:69-75 melspectrogram: preemphasis :22-25, _stft :139-143, _linear_to_mel :181-185,
_build_mel_basis :193-199, _amp_to_db :201-203, _normalize :208-212; hparams.py:18-34).

librosa is not installable here, so librosa.stft / librosa.filters.mel are restated from their published
algorithms (librosa 0.6-0.7, the versions the reference's API use implies):
  * stft(n_fft=2048, hop_length=300, win_length=1200): center=True with reflect padding of n_fft//2,
    periodic Hann window of win_length zero-padded on both sides to n_fft, frames = 1 + len(y)//hop,
    FFT evaluated in float64 (y is float64 after scipy.signal.lfilter), result stored as complex64.
  * filters.mel(sr, n_fft, n_mels): Slaney mel scale (linear below 1 kHz, log above), triangular
    filters, Slaney area normalisation, fmin=0, fmax=sr/2, float32.
PARITY: the reference's own utils/audio.py + hparams.py, run unmodified with librosa stubbed by the functions below
(tests/golden/make_reference_audio_golden.py -> ref_audio.npz), is reproduced to 1e-5 (glue, constants, clipping, dB floor);
librosa itself is UNPINNED: stft / filters.mel are pinned against torch.stft and torchaudio.functional.melscale_fbanks
(tests/test_mel.py), which document librosa compatibility.
"""
import numpy as np
from scipy import signal

DEFAULTS = dict(sample_rate=24000, fft_size=2048, hop_size=300, win_size=1200, num_mels=80, preemphasis=0.97,
                preemphasize=True, min_level_db=-100, ref_level_db=20, max_abs_value=4.0)


def preemphasis(wav, k, preemphasize=True):
    # utils/audio.py:22-25
    if preemphasize:
        return signal.lfilter([1, -k], [1], wav)
    return wav


def hann_padded(win_length, n_fft):
    w = signal.get_window('hann', win_length, fftbins=True)
    lpad = (n_fft - win_length) // 2
    return np.pad(w, (lpad, n_fft - win_length - lpad))


def stft(y, n_fft, hop, win_length):
    y = np.asarray(y, dtype=np.float64)
    w = hann_padded(win_length, n_fft)
    ypad = np.pad(y, n_fft // 2, mode='reflect')
    n_frames = 1 + (len(ypad) - n_fft) // hop
    frames = np.stack([ypad[t * hop:t * hop + n_fft] for t in range(n_frames)], axis=1)
    return np.fft.rfft(w[:, None] * frames, axis=0).astype(np.complex64)


def hz_to_mel(f):
    f = np.asarray(f, dtype=np.float64)
    f_sp = 200.0 / 3
    mels = f / f_sp
    min_log_hz, min_log_mel, logstep = 1000.0, 1000.0 / f_sp, np.log(6.4) / 27.0
    return np.where(f >= min_log_hz, min_log_mel + np.log(np.maximum(f, 1e-10) / min_log_hz) / logstep, mels)


def mel_to_hz(m):
    m = np.asarray(m, dtype=np.float64)
    f_sp = 200.0 / 3
    min_log_hz, min_log_mel, logstep = 1000.0, 1000.0 / f_sp, np.log(6.4) / 27.0
    return np.where(m >= min_log_mel, min_log_hz * np.exp(logstep * (m - min_log_mel)), f_sp * m)


def mel_basis(sr, n_fft, n_mels, fmin=0.0, fmax=None):
    fmax = sr / 2.0 if fmax is None else fmax
    fftfreqs = np.linspace(0, sr / 2.0, 1 + n_fft // 2)
    mel_f = mel_to_hz(np.linspace(hz_to_mel(fmin), hz_to_mel(fmax), n_mels + 2))
    fdiff = np.diff(mel_f)
    ramps = mel_f[:, None] - fftfreqs[None, :]
    weights = np.zeros((n_mels, 1 + n_fft // 2))
    for i in range(n_mels):
        lower = -ramps[i] / fdiff[i]
        upper = ramps[i + 2] / fdiff[i + 1]
        weights[i] = np.maximum(0, np.minimum(lower, upper))
    enorm = 2.0 / (mel_f[2:n_mels + 2] - mel_f[:n_mels])
    weights *= enorm[:, None]
    return weights.astype(np.float32)


def melspectrogram(wav, **hp):
    """utils/audio.py:69-75.  wav float32 1-D -> (num_mels, frames) float32, like the reference."""
    h = dict(DEFAULTS); h.update(hp)
    y = preemphasis(np.asarray(wav, np.float32), h['preemphasis'], h['preemphasize'])
    D = stft(y, h['fft_size'], h['hop_size'], h['win_size'])
    mag = np.abs(D)                                                       # float32
    mel = np.dot(mel_basis(h['sample_rate'], h['fft_size'], h['num_mels']), mag)      # float32
    min_level = np.exp(h['min_level_db'] / 20 * np.log(10))
    S = 20 * np.log10(np.maximum(min_level, mel)) - h['ref_level_db']     # :201-203, :71
    S = S.astype(np.float32)
    m = h['max_abs_value']
    return np.clip((2 * m) * ((S - h['min_level_db']) / (-h['min_level_db'])) - m, -m, m).astype(np.float32)


def synthetic_speech(n, sr=24000, seed=0):
    """Deterministic speech-like test signal: a few harmonics with a pitch glide and a syllabic envelope,
    plus a little noise and a silent gap (exercises the dB floor)."""
    rs = np.random.RandomState(seed)
    t = np.arange(n) / sr
    f0 = 120 + 40 * np.sin(2 * np.pi * 0.7 * t)
    phase = 2 * np.pi * np.cumsum(f0) / sr
    x = sum((0.6 / k) * np.sin(k * phase + rs.rand() * 6.28) for k in range(1, 24))
    env = (0.5 + 0.5 * np.sin(2 * np.pi * 3.1 * t)) ** 2
    x = x * env + 0.003 * rs.randn(n)
    gap = slice(n // 3, n // 3 + sr // 10)
    x[gap] = 0.0
    return (0.5 * x / np.max(np.abs(x))).astype(np.float32)
