"""One bf16 training step through the tcgen05 kernels at the 30-layer R=D=128 sizes on a short crop (compute-sanitizer target)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tacotron_wavenet_vocoder_korean_b200 import synth                                  # noqa: E402
from tacotron_wavenet_vocoder_korean_b200.wavenet.train import WaveNetTrainer           # noqa: E402
from tests.train_helpers import train_case                                               # noqa: E402

kw = dict(synth.cfg2(2), dilations=[1, 2, 4, 8, 16, 32, 64, 128, 256, 512])
T = 1500
w, wav, mel, gc = train_case(kw, T)
tr = WaveNetTrainer(T, dtype='bf16', **kw)
tr.load_state_dict(w)
hp = dict(wavenet_learning_rate=1e-3, wavenet_decay_rate=0.5, wavenet_decay_steps=300000, wavenet_clip_gradients=True)
for _ in range(2):
    loss = tr.train_step(wav, mel, gc, hp)
print('loss', float(loss.item()), tr.info()['fused_launches'])
