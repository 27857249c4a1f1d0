"""Throughput of the STFT->mel kernel (next-2): audio seconds per second, vs the numpy oracle on the host."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import mel_oracle as mo
from tacotron_wavenet_vocoder_korean_b200 import audio
from tacotron_wavenet_vocoder_korean_b200.hparams import hparams
rows, secs = 64, 10
x = np.stack([mo.synthetic_speech(24000 * secs, seed=s % 4) for s in range(rows)])
xd = torch.from_numpy(x).cuda()
for _ in range(3):
    m = audio.melspectrogram(xd, hparams)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    m = audio.melspectrogram(xd, hparams)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
t0 = time.perf_counter(); ref = mo.melspectrogram(x[0]); cpu_s = time.perf_counter() - t0
frames = m.shape[2]
alg_bytes = x.nbytes + m.numel() * 4
print(json.dumps({"workload": "melspectrogram %d x %d s @ 24 kHz (n_fft 2048, hop 300, win 1200, 80 mels)" % (rows, secs),
                  "gpu_ms": ms, "audio_seconds_per_second": rows * secs / (ms / 1e3), "frames_per_second": rows * frames / (ms / 1e3),
                  "algorithmic_bytes": alg_bytes, "achieved_GBps": alg_bytes / (ms / 1e3) / 1e9,
                  "cpu_oracle_audio_seconds_per_second": secs / cpu_s, "max_abs_diff_row0": float(np.abs(m[0].cpu().numpy() - ref).max())}))
