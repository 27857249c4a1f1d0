import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tests.helpers import make_inputs
from tacotron_wavenet_vocoder_korean_b200 import synth, _lib
from tacotron_wavenet_vocoder_korean_b200.wavenet import WaveNetModel
lib = _lib.lib()
lib.wn_debug_profile.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int]
N = int(sys.argv[1]) if len(sys.argv) > 1 else 16
T = 2000
kw = synth.cfg2(N); w = synth.make_weights(**kw)
net = WaveNetModel(train_mode=False, **kw); net.load_state_dict(w)
inp = make_inputs(kw, T); lc = net.create_upsample(inp['mel'])
net.generate(T, inp['x0'], inp['uniforms'], lc_up=lc, gc_ids=inp['gc_ids'])
lib.wn_debug_profile(net._h, 1, None, 0)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); net.generate(T, inp['x0'], inp['uniforms'], lc_up=lc, gc_ids=inp['gc_ids'], sync=False); e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
grid = net.info()['grid']
buf = np.zeros(grid * 16, np.int64); lib.wn_debug_profile(net._h, 0, buf.ctypes.data_as(C.c_void_p), buf.size)
p = buf.reshape(grid, 16) / float(T * N)
print('N=%d %.1f us/step %.0f samples/s; cycles per row-step = %.0f' % (N, 1e3 * ms / T, N * T / ms * 1e3, ms * 1e-3 * 1.965e9 / T / N))
names = ['wait_x', 'fg+act', 'dense', 'zgather', 'skip+acc', 'pre']
L, M, Mt = 30, 4, 16
for l in (0, 1, 2, 15, 28, 29):
    print(' layer %2d:' % l, ' '.join('%s=%d' % (n, p[l * M:(l + 1) * M, i].mean()) for i, n in enumerate(names)), ' total=%d' % p[l * M:(l + 1) * M, :6].sum(axis=1).mean())
print(' tail   :', ' '.join('%s=%d' % (n, p[L * M:L * M + Mt, i].mean()) for i, n in enumerate(['wait_acc', 'post1', 'post2'])))
print(' sampler:', ' '.join('%s=%d' % (n, p[-1, i]) for i, n in enumerate(['wait_c2', 'draw', 'feed'])))
