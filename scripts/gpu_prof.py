"""Phase profile of the persistent kernel (diagnostic): average cycles per (row, step) in each phase."""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tests.helpers import make_inputs
from tacotron_wavenet_vocoder_korean_b200 import synth, _lib
from tacotron_wavenet_vocoder_korean_b200.wavenet import WaveNetModel

lib = _lib.lib()
lib.wn_debug_pingpong.restype = C.c_longlong
lib.wn_debug_pingpong.argtypes = [C.c_int]
lib.wn_debug_profile.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int]
print('LL ping-pong round trip (cycles):', lib.wn_debug_pingpong(20000), lib.wn_debug_pingpong(20000))


def prof(name, kw, T):
    w = synth.make_weights(**kw)
    net = WaveNetModel(train_mode=False, **kw); net.load_state_dict(w)
    inp = make_inputs(kw, T)
    lc = net.create_upsample(inp['mel']) if 'mel' in inp else None
    info = net.info(); grid = info['grid']; N = kw['batch_size']
    net.generate(T, inp['x0'], inp['uniforms'], lc_up=lc, gc_ids=inp['gc_ids'])
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); net.generate(T, inp['x0'], inp['uniforms'], lc_up=lc, gc_ids=inp['gc_ids'], sync=False); e1.record(); torch.cuda.synchronize()
    print('%s N=%d unprofiled: %.2f us/step (%.0f samples/s)' % (name, N, 1e3 * e0.elapsed_time(e1) / T, N * T / e0.elapsed_time(e1) * 1e3))
    lib.wn_debug_profile(net._h, 1, None, 0)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); net.generate(T, inp['x0'], inp['uniforms'], lc_up=lc, gc_ids=inp['gc_ids'], sync=False); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    buf = np.zeros(grid * 16, np.int64)
    lib.wn_debug_profile(net._h, 0, buf.ctypes.data_as(C.c_void_p), buf.size)
    p = buf.reshape(grid, 16) / float(T * N)
    L, M, Mt = len(kw['dilations']), info['M'], info['Mt']
    print('%s: N=%d T=%d  %.1f us/step  (%.0f samples/s)' % (name, N, T, 1e3 * ms / T, N * T / ms * 1e3))
    lay = p[:L * M].reshape(L, M, 16)
    names = ['wait_x', 'fg+act', 'dense', 'zgather', 'skip+acc', 'pre']
    print('  layer phases (cycles per row-step), mean over CTAs:', {n: int(lay[:, :, i].mean()) for i, n in enumerate(names)})
    for l in (0, 1, L // 2, L - 1):
        print('   layer %2d m0:' % l, {n: int(lay[l, 0, i]) for i, n in enumerate(names)})
    tl = p[L * M:L * M + Mt]
    print('  tail: ', {n: int(tl[:, i].mean()) for i, n in enumerate(['wait_acc', 'post1', 'post2'])})
    print('  sampler:', {n: int(p[-1, i]) for i, n in enumerate(['wait_c2', 'draw', 'feed'])})
    busy = lay[:, :, 1:6].sum(axis=2).mean()
    print('  extra marks: [6]=wait until own words arrive, phase0=barrier after poll; [7]=dense dot, phase2=epilogue+post:', int(lay[1:, :, 6].mean()), int(lay[1:, :, 0].mean()), int(lay[:-1, :, 7].mean()), int(lay[:-1, :, 2].mean()))
    print('  layer busy cycles per row-step %.0f; chain part (fg+dense) %.0f' % (busy, lay[:, :, 1:3].sum(axis=2).mean()))


import time
prof("cfg2", synth.cfg2(1), 3000)
net = WaveNetModel(train_mode=False, **synth.cfg2(8)); net.load_state_dict(synth.make_weights(**synth.cfg2(8))); print('info', net.info())
