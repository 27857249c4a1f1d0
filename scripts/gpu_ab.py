"""A/B timing of kernel variants inside one process (same box, same clocks)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tests.helpers import make_inputs
from tacotron_wavenet_vocoder_korean_b200 import synth
from tacotron_wavenet_vocoder_korean_b200.wavenet import WaveNetModel

def timeit(kw, T, reps=3, **extra):
    net = WaveNetModel(train_mode=False, **kw, **extra); net.load_state_dict(synth.make_weights(**kw))
    inp = make_inputs(kw, T)
    lc = net.create_upsample(inp['mel']) if 'mel' in inp else None
    net.generate(T, inp['x0'], inp['uniforms'], lc_up=lc, gc_ids=inp['gc_ids'])
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); net.generate(T, inp['x0'], inp['uniforms'], lc_up=lc, gc_ids=inp['gc_ids'], sync=False); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return 1e3 * best / T, net.info()

for name, kw, T in [('cfg2 N=1', synth.cfg2(1), 4000), ('cfg2 N=8', synth.cfg2(8), 4000), ('cfg1 N=1', synth.cfg1(1), 6000), ('hparams N=2', synth.cfg_hparams_default(2), 3000)]:
    for label, extra in [('die-aware', {}), ('single-homed', {'die_aware': False}), ('generic kernel', {'generic_kernel': True})]:
        us, info = timeit(kw, T, **extra)
        print('%-12s %-15s %.2f us/step  %.0f samples/s  (die_aware=%d static=%d)' % (name, label, us, kw['batch_size'] * 1e6 / us, info['die_aware'], info['static_shape']), flush=True)
