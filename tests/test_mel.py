"""STFT -> mel chain (reference utils/audio.py:69-75).
CPU: the numpy oracle against torch.stft / torchaudio (documented librosa-compatible) and its own invariants.
GPU: the sm_100a kernel against the oracle within 1e-4 (north_star's tolerance for float mel)."""
import numpy as np
import pytest
import torch

from oracle import mel_oracle as mo
from tacotron_wavenet_vocoder_korean_b200.hparams import hparams


def test_stft_matches_torch_stft():
    x = mo.synthetic_speech(24000 + 77, seed=1)
    y = mo.preemphasis(x, 0.97)
    D = mo.stft(y, 2048, 300, 1200)
    w = torch.from_numpy(mo.signal.get_window('hann', 1200, fftbins=True))
    Dt = torch.stft(torch.from_numpy(y), n_fft=2048, hop_length=300, win_length=1200, window=w, center=True,
                    pad_mode='reflect', return_complex=True).numpy()
    assert D.shape == Dt.shape == (1025, 1 + len(x) // 300)
    assert np.abs(D - Dt).max() < 1e-6 * np.abs(D).max()


def test_mel_basis_matches_torchaudio_slaney():
    torchaudio = pytest.importorskip('torchaudio')
    fb = torchaudio.functional.melscale_fbanks(n_freqs=1025, f_min=0.0, f_max=12000.0, n_mels=80, sample_rate=24000,
                                               norm='slaney', mel_scale='slaney').numpy().T
    mb = mo.mel_basis(24000, 2048, 80)
    assert mb.shape == (80, 1025) and mb.dtype == np.float32
    assert np.abs(fb - mb).max() < 2e-7
    assert np.all(mb >= 0) and np.all((mb > 0).sum(axis=1) >= 2)


def test_oracle_melspectrogram_invariants():
    x = mo.synthetic_speech(24000 * 2, seed=2)
    M = mo.melspectrogram(x)
    assert M.shape == (80, 161) and M.dtype == np.float32                  # frames = 1 + len // hop
    assert M.min() >= -4.0 and M.max() <= 4.0                               # _normalize clip, audio.py:208-212
    assert np.all(mo.melspectrogram(np.zeros(5000, np.float32)) == -4.0)    # silence sits on the dB floor
    # datasets/moon.py:113-146: audio is cut to frames * hop; frames of a hop-multiple signal
    assert mo.melspectrogram(x[:300 * 40]).shape[1] == 41
    # pre-emphasis is y[n] = x[n] - 0.97 x[n-1] (audio.py:22-25)
    y = mo.preemphasis(x, 0.97)
    assert y.dtype == np.float64 and abs(y[5] - (np.float64(x[5]) - 0.97 * np.float64(x[4]))) < 1e-15


@pytest.mark.gpu
@pytest.mark.parametrize('n,seed', [(24000 * 5, 0), (24000 * 2 + 131, 3), (3000, 4), (1025, 5)])
def test_gpu_melspectrogram_matches_oracle(n, seed):
    from tacotron_wavenet_vocoder_korean_b200 import audio
    x = mo.synthetic_speech(n, seed=seed)
    ref = mo.melspectrogram(x)
    got = audio.melspectrogram(torch.from_numpy(x), hparams).cpu().numpy()
    assert got.shape == ref.shape and got.dtype == np.float32
    assert np.abs(got - ref).max() < 1e-4, np.abs(got - ref).max()          # float mel within 1e-4 (north_star)


@pytest.mark.gpu
def test_gpu_melspectrogram_batch_edges_errors():
    from tacotron_wavenet_vocoder_korean_b200 import audio
    xs = np.stack([mo.synthetic_speech(7200, seed=s) for s in range(3)])
    got = audio.melspectrogram(torch.from_numpy(xs), hparams).cpu().numpy()
    assert got.shape == (3, 80, 25)
    for r in range(3):
        assert np.abs(got[r] - mo.melspectrogram(xs[r])).max() < 1e-4
    assert torch.all(audio.melspectrogram(torch.zeros(4000), hparams) == -4.0)
    loud = audio.melspectrogram(torch.from_numpy((30 * xs[0]).astype(np.float32)), hparams)
    assert float(loud.max()) <= 4.0
    with pytest.raises(RuntimeError):
        audio.melspectrogram(torch.zeros(500), hparams)                      # shorter than fft_size / 2: cannot reflect-pad


@pytest.mark.gpu
@pytest.mark.parametrize('fft_size,win_size,hop', [(2048, 1200, 300), (512, 400, 100), (1024, 1024, 256), (4096, 2400, 600), (8192, 4800, 1200), (64, 64, 16)])
def test_gpu_melspectrogram_fft_sizes_and_radix_paths(fft_size, win_size, hop):
    """Every factorisation the Stockham kernel takes (8*8*8*4, 8*8*8, 8*8*8*2, 8^4, 8*8) and the radix-2 kernel (n_fft 8192, and
    WN_MEL_RADIX2 for the default size) against the oracle within 1e-4; odd frame counts leave the last pair half empty."""
    import copy
    import os
    from tacotron_wavenet_vocoder_korean_b200 import audio
    hp = copy.copy(hparams)
    hp.fft_size, hp.win_size, hp.hop_size = fft_size, win_size, hop
    if fft_size == 64:
        hp.num_mels = 8
    for n in (24000 + 77, hop * 20):                       # 1 + n // hop frames: even and odd counts
        x = mo.synthetic_speech(n, seed=1)
        ref = mo.melspectrogram(x, fft_size=fft_size, win_size=win_size, hop_size=hop, num_mels=hp.num_mels)
        got = audio.melspectrogram(torch.from_numpy(x), hp).cpu().numpy()
        assert got.shape == ref.shape
        assert np.abs(got - ref).max() < 1e-4, np.abs(got - ref).max()
