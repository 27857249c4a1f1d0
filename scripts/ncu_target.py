"""Short cfg2 run for Nsight Compute captures (never used for reported numbers)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tests.helpers import make_inputs
from tacotron_wavenet_vocoder_korean_b200 import synth
from tacotron_wavenet_vocoder_korean_b200.wavenet import WaveNetModel

T = int(sys.argv[1]) if len(sys.argv) > 1 else 3000
kw = synth.cfg2(8)
net = WaveNetModel(train_mode=False, **kw)
net.load_state_dict(synth.make_weights(**kw))
inp = make_inputs(kw, T)
lc = net.create_upsample(inp['mel'])
for _ in range(2):
    net.generate(T, inp['x0'], inp['uniforms'], lc_up=lc, gc_ids=inp['gc_ids'])
torch.cuda.synchronize()
print('done', net.info())
