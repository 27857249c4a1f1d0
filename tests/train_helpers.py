"""Seeded inputs for the training-step parity tests (same arrays go to oracle/train_oracle.py and to the CUDA path)."""
import numpy as np

from tacotron_wavenet_vocoder_korean_b200 import synth


def train_case(kw, T, seed=11, weight_seed=1234):
    """-> weights, wav (N,T), mel (N,T/hop,C) or None, gc ids (N) or None."""
    rs = np.random.RandomState(seed)
    N = kw['batch_size']
    w = synth.make_weights(seed=weight_seed, **kw)
    t = np.arange(T)[None, :]
    wav = 0.6 * np.sin(2 * np.pi * t * rs.uniform(0.01, 0.05, (N, 1)) + rs.uniform(0, 6, (N, 1))) + 0.15 * rs.randn(N, T)
    wav = np.clip(wav, -1, 1).astype(np.float32)
    wav[0, T - 3] = -1.0           # exercise the y < -0.999 / y > 0.999 branches of the loss
    wav[N - 1, T - 2] = 1.0
    mel = gc = None
    if kw.get('local_condition_channels'):
        hop = int(np.prod(kw['upsample_factor']))
        assert T % hop == 0
        mel = np.clip(rs.randn(N, T // hop, kw['local_condition_channels']) * 1.5, -4, 4).astype(np.float32)
    if kw.get('global_condition_channels'):
        gc = (np.arange(N) % kw['global_condition_cardinality']).astype(np.int32)
    return w, wav, mel, gc


def at_cell_centres(wav, quantization_channels=256):
    """Snap a waveform to the centres of its mu-law cells.  The one-hot model encodes the waveform inside the step
    (wavenet/model.py:257, fp32); at a centre every libm / libdevice log1p yields the same id, so the oracle and the CUDA
    path see the same one-hot input."""
    from oracle import np_oracle
    ids = np_oracle.mu_law_encode(wav, quantization_channels)
    return np.asarray(np_oracle.mu_law_decode(ids, quantization_channels), np.float32)


def rel_err(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-12))


def cosine(a, b):
    a, b = np.asarray(a, np.float64).ravel(), np.asarray(b, np.float64).ravel()
    return float(a @ b / max(np.linalg.norm(a) * np.linalg.norm(b), 1e-300))


def well_conditioned(w):
    """Wide logistic components (log-scale bias -0.5) and small output weights.  With synth's default head (log-scale bias -3)
    every output gradient is amplified by 1/scale ~ 20-50 and bf16 rounding of the logits alone moves the gradients by tens
    of percent; this variant makes a bf16 step comparable with the fp32 oracle at the per-cent level."""
    w = dict(w)
    w['wavenet/conv1d_2/kernel'] = w['wavenet/conv1d_2/kernel'] * 0.2
    b = w['wavenet/conv1d_2/bias'].copy()
    k = b.shape[0] // 3
    b[2 * k:] = -0.5
    w['wavenet/conv1d_2/bias'] = b
    return w


# ---- a small on-disk data set in the format preprocess.py writes (datasets/moon.py:158-172), for the feeder pin ----------------
FEEDER_HP = dict(sample_size=100, hop_size=12, num_mels=5, skip_path_filter=False)


def make_feeder_dataset(root, dirs=('spk_a', 'spk_b'), n_files=(7, 5), seed=5):
    """Deterministic npz files + train.txt under root/<dir>: audio (frames*hop,), mel (frames, num_mels), time_steps, mel_frames.
    Some utterances are shorter than sample_size (filtered by train.txt), one listed file is missing on disk."""
    import os
    rs = np.random.RandomState(seed)
    hop, C = FEEDER_HP['hop_size'], FEEDER_HP['num_mels']
    for di, (d, n) in enumerate(zip(dirs, n_files)):
        os.makedirs(os.path.join(root, d), exist_ok=True)
        lines = []
        for i in range(n):
            frames = int(rs.randint(5, 30))
            name = 'utt%d_%02d.npz' % (di, i)
            audio = (rs.rand(frames * hop) * 2 - 1).astype(np.float32)
            mel = rs.randn(frames, C).astype(np.float32)
            if not (di == 0 and i == 3):                   # listed in train.txt but absent on disk (datafeeder_wavenet.py:132-135)
                np.savez(os.path.join(root, d, name), audio=audio, mel=mel, time_steps=frames * hop, mel_frames=frames)
            lines.append('|'.join(['a', 'm', 'l', str(frames * hop), str(frames), 'text', name]))
        with open(os.path.join(root, d, 'train.txt'), 'w', encoding='utf-8') as f:
            f.write('\n'.join(lines) + '\n')
