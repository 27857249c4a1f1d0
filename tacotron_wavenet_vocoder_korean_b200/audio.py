"""The one audio helper on the WaveNet path: utils/audio.py:14-17 save_wav."""
import numpy as np
from scipy.io import wavfile


def save_wav(wav, path, sr):
    """Peak-normalise to int16 and write (utils/audio.py:14-17).  Unlike the reference this does not
    scale the caller's array in place."""
    wav = np.asarray(wav, dtype=np.float64)
    wav = wav * (32767 / max(0.01, np.max(np.abs(wav)) if wav.size else 0.0))
    wavfile.write(path, sr, wav.astype(np.int16))
