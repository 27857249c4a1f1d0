"""Debug aid: per-variable error table of the CUDA training step against the torch oracle (run under gpurun)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import train_oracle as to                      # noqa: E402
from tacotron_wavenet_vocoder_korean_b200 import synth     # noqa: E402
from tacotron_wavenet_vocoder_korean_b200.wavenet.train import WaveNetTrainer   # noqa: E402
from tests.train_helpers import train_case, rel_err, cosine   # noqa: E402

for dtype in ('fp32', 'bf16'):
    kw = synth.tiny_train(3)
    T = 96
    w, wav, mel, gc = train_case(kw, T)
    om = to.TorchWaveNetTrain(w, **kw)
    Lo, go = om.loss_and_grads(wav, mel, gc, 0.01)
    tr = WaveNetTrainer(T, dtype=dtype, **kw)
    tr.load_state_dict(w)
    L = float(tr.loss_and_grads(wav, mel, gc, 0.01).item())
    raw_o = om.raw_output(wav, mel, gc)[0].detach().numpy().reshape(-1, 30)
    raw = tr.debug_get('raw_output').reshape(-1, 30)
    print("== %s: loss %.6f oracle %.6f   raw_output max err %.3g" % (dtype, L, Lo, np.abs(raw - raw_o).max()))
    lc_o = om.create_upsample(__import__('torch').from_numpy(mel)).detach().numpy()
    lc = tr.debug_get('lc').reshape(3, -1, kw['local_condition_channels'])
    print("   lc max err %.3g" % np.abs(lc - lc_o[:, :lc.shape[1]]).max())
    g = tr.state_dict('grads')
    for k in go:
        print("   %-70s rel %.3g cos %.5f |g| %.3g" % (k, rel_err(g[k], go[k]), cosine(g[k], go[k]), np.linalg.norm(go[k])))
print(tr.info())
