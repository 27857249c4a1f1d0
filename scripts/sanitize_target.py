import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tests.helpers import make_inputs
from tacotron_wavenet_vocoder_korean_b200 import synth
from tacotron_wavenet_vocoder_korean_b200.wavenet import WaveNetModel
for name, kw, T in [('tiny_mol', synth.tiny_mol(), 12), ('tiny_mulaw', synth.tiny_mulaw(), 12), ('cfg2 n2', synth.cfg2(2), 6), ('cfg2 n10 (ws)', synth.cfg2(10), 4), ('cfg1', synth.cfg1(), 6)]:
    net = WaveNetModel(train_mode=False, **kw); net.load_state_dict(synth.make_weights(**kw))
    inp = make_inputs(kw, T)
    lc = net.create_upsample(inp['mel']) if 'mel' in inp else None
    s = net.generate(T, inp['x0'], inp['uniforms'], lc_up=lc, gc_ids=inp['gc_ids'])
    torch.cuda.synchronize()
    print(name, 'ok', float(s.abs().sum()))
