"""r02 A/B: cluster / DSMEM layer kernel (wn_kernel_v2.cuh) against the round-1 single-kernel path, same process.
Checks bit-identical samples + logits, then times both at 1 / 8 / 16 rows and prints the in-kernel phase counters."""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tests.helpers import make_inputs
from tacotron_wavenet_vocoder_korean_b200 import synth, _lib
from tacotron_wavenet_vocoder_korean_b200.wavenet import WaveNetModel

lib = _lib.lib()
lib.wn_debug_profile.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int]


def build(kw, no16=False, **extra):
    if no16:
        os.environ['WN_NO_CLUSTER16'] = '1'
    net = WaveNetModel(train_mode=False, **kw, **extra)
    net.load_state_dict(synth.make_weights(**kw))
    os.environ.pop('WN_NO_CLUSTER16', None)
    return net


def timed(net, T, inp, lc, rows):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    args = (T, inp['x0'][:rows], inp['uniforms'][:rows])
    kw = dict(lc_up=lc[:rows], gc_ids=inp['gc_ids'][:rows])
    net.generate(*args, **kw)
    e0.record(); net.generate(*args, sync=False, **kw); e1.record(); torch.cuda.synchronize()
    net.sync_check()
    return e0.elapsed_time(e1)


def main():
    mode = sys.argv[1] if len(sys.argv) > 1 else 'all'
    T = 400
    kw = synth.cfg2(8)
    inp = make_inputs(kw, T)
    a = build(kw)
    b = build(kw, cluster=False)
    print('info v2:', a.info())
    print('info v1:', b.info())
    lc = a.create_upsample(inp['mel'])
    for rows in (1, 2, 8):
        sa, la = a.generate(T, inp['x0'][:rows], inp['uniforms'][:rows], lc_up=lc[:rows], gc_ids=inp['gc_ids'][:rows], want_logits=True)
        sb, lb = b.generate(T, inp['x0'][:rows], inp['uniforms'][:rows], lc_up=lc[:rows], gc_ids=inp['gc_ids'][:rows], want_logits=True)
        same = torch.equal(sa, sb) and torch.equal(la, lb)
        print('rows=%d T=%d v2 == v1 bit-exact: %s' % (rows, T, same))
        if not same:
            d = (la != lb).nonzero()
            print('  first logit mismatch at', d[0].tolist() if len(d) else None, 'max |diff|', float((la - lb).abs().max()))
            d = (sa != sb).nonzero()
            print('  first sample mismatch at', d[0].tolist() if len(d) else None)
    # ragged rows
    T_row = [400, 17, 0, 255, 400, 1, 399, 64]
    sa = a.generate(T, inp['x0'], inp['uniforms'], lc_up=lc, gc_ids=inp['gc_ids'], T_row=T_row)
    sb = b.generate(T, inp['x0'], inp['uniforms'], lc_up=lc, gc_ids=inp['gc_ids'], T_row=T_row)
    print('ragged rows equal:', all(torch.equal(sa[r, :t], sb[r, :t]) for r, t in enumerate(T_row)))
    # folded create_upsample (mel frames staged by TMA) against the materialised upsampled condition
    T2 = 1300
    inp2 = make_inputs(kw, T2)
    lc2 = a.create_upsample(inp2['mel'])
    s_lc = a.generate(T2, inp2['x0'], inp2['uniforms'], lc_up=lc2, gc_ids=inp2['gc_ids'], want_logits=True)
    s_mel = a.generate(T2, inp2['x0'], inp2['uniforms'], mel=inp2['mel'], gc_ids=inp2['gc_ids'], want_logits=True)
    print('mel-folded == materialised lc: %s' % (torch.equal(s_lc[0], s_mel[0]) and torch.equal(s_lc[1], s_mel[1])))
    s_lc = a.generate(T2, inp2['forced_full'][:, :400], inp2['uniforms'], lc_up=lc2, lc_shift=399, gc_ids=inp2['gc_ids'])
    s_mel = a.generate(T2, inp2['forced_full'][:, :400], inp2['uniforms'], mel=inp2['mel'], lc_shift=399, gc_ids=inp2['gc_ids'])
    print('mel-folded == materialised lc with priming (lc_shift 399): %s' % torch.equal(s_lc, s_mel))
    s_v1 = b.generate(T2, inp2['x0'], inp2['uniforms'], mel=inp2['mel'], gc_ids=inp2['gc_ids'])
    print('single-kernel path with mel argument (materialises internally) equal: %s' % torch.equal(s_v1, a.generate(T2, inp2['x0'], inp2['uniforms'], mel=inp2['mel'], gc_ids=inp2['gc_ids'])))
    if mode == 'check':
        return
    if mode == 'time':
        T = 2000
        kw16 = synth.cfg2(16)
        inp = make_inputs(kw16, T)
        a16 = build(kw16)
        c8 = build(kw16, no16=True)
        print('info:', a16.info()['cluster_path'], c8.info()['cluster_path'])
        lc = a16.create_upsample(inp['mel'])
        for rows in (1, 8, 12, 16):
            for name, net in (('v2 16+8 clusters', a16), ('v2 8-clusters   ', c8)):
                ms = min(timed(net, T, inp, lc, rows) for _ in range(2))
                print('%s rows=%2d: %.2f us/step  %.1f k samples/s' % (name, rows, 1e3 * ms / T, rows * T / ms))
        return
    T = 3000
    kw16 = synth.cfg2(16)
    inp = make_inputs(kw16, T)
    a16, b16 = build(kw16), build(kw16, cluster=False)
    lc = a16.create_upsample(inp['mel'])
    f16 = build(kw16, fast_act=True)
    print('info fast:', f16.info())
    # fast activation: teacher-forced logits against the pinned arithmetic (tolerance of north_star: 1e-4)
    Tt = 1200
    fa = f16.generate(Tt, inp['forced_full'][:4, :Tt], inp['uniforms'][:4, :Tt], lc_up=lc[:4], gc_ids=inp['gc_ids'][:4], want_logits=True)
    ex = a16.generate(Tt, inp['forced_full'][:4, :Tt], inp['uniforms'][:4, :Tt], lc_up=lc[:4], gc_ids=inp['gc_ids'][:4], want_logits=True)
    print('fast vs pinned activation, teacher forced %d steps: max |dlogit| %.3g, max |dsample| %.3g' %
          (Tt, float((fa[1] - ex[1]).abs().max()), float((fa[0] - ex[0]).abs().max())))
    for rows in (1, 2, 4, 8, 12, 16):
        for name, net in (('v2', a16), ('v2fast', f16), ('v1', b16)):
            ms = min(timed(net, T, inp, lc, rows) for _ in range(2))
            print('%s rows=%2d: %.2f us/step  %.1f k samples/s' % (name, rows, 1e3 * ms / T, rows * T / ms))
    # phase counters of the v2 path, 1 and 8 rows
    for rows in (1, 8, 16):
        lib.wn_debug_profile(a16._h, 1, None, 0)
        a16.generate(T, inp['x0'][:rows], inp['uniforms'][:rows], lc_up=lc[:rows], gc_ids=inp['gc_ids'][:rows])
        grid = a16.info()['grid']
        buf = np.zeros(grid * 16, np.int64)
        lib.wn_debug_profile(a16._h, 0, buf.ctypes.data_as(C.c_void_p), buf.size)
        raw = buf.reshape(grid, 16).astype(np.float64)
        p = raw / float(T * rows)
        L = 30
        lay = p[:L * 4].reshape(L, 4, 16)
        names = {0: 'wait', 1: 'combine', 2: 'bar1', 3: 'lds+fma+reduce', 4: 'act+zsend', 5: 'bar2', 6: 'dense+send',
                 7: 'h:wait_full', 8: 'h:ring+zwait', 9: 'h:skip+acc', 10: 'h:pre'}
        print('rows=%d layer phases (cycles per row-step)' % rows)
        for sel, tag in ((slice(2, L, 2), 'even layers (LL in, DSMEM out)'), (slice(1, L, 2), 'odd layers (DSMEM in, LL out)')):
            print('   %s:' % tag, {n: int(lay[sel, :, i].mean()) for i, n in names.items()})
        print('   layer 0:', {n: int(lay[0, :, i].mean()) for i, n in names.items()})
        print('   layer 29:', {n: int(lay[29, :, i].mean()) for i, n in names.items()})
        tl = p[L * 4:L * 4 + 16]
        print('   tail:', {n: int(tl[:, i].mean()) for i, n in enumerate(['wait_acc', 'post1', 'post2'])})
        print('   per-layer chain wait :', [int(v) for v in lay[:, :, 0].mean(axis=1)])
        print('   per-layer chain busy :', [int(v) for v in lay[:, :, 1:7].sum(axis=2).mean(axis=1)])
        print('   per-layer helper wait:', [int(v) for v in lay[:, :, 7].mean(axis=1)])
        print('   per-layer helper busy:', [int(v) for v in lay[:, :, 8:11].sum(axis=2).mean(axis=1)])
        print('   layer 0 sampler (chain group, per row-step): noise / forced input from shared %d, poll of the tail partials (4 words per lane per warp) %d, '
              'barrier %d; partial sum + draw + causal FMA are in combine' % (lay[0, :, 11].mean(), lay[0, :, 12].mean(), lay[0, :, 13].mean()))
        print('   tail wait / busy per CTA:', [int(v) for v in tl[:, 0]], [int(v) for v in tl[:, 1:3].sum(axis=1)])
        if rows >= 1:
            # timeline from the global-timer stamps (ns -> cycles at 1.965 GHz): wake = input complete, send = outputs posted
            ghz = 1.965
            wake = lay[:, :, 14].mean(axis=1) * ghz
            send = lay[:, :, 15].mean(axis=1) * ghz
            hop = wake[1:] - send[:-1]
            comp = send - wake
            print('   hop cycles into odd layers (DSMEM):', [int(v) for v in hop[0::2]])
            print('   hop cycles into even layers (LL):  ', [int(v) for v in hop[1::2]])
            print('   wake->send cycles per layer:', [int(v) for v in comp])
            print('   means: DSMEM hop %.0f, LL hop %.0f, layer wake->send %.0f' % (hop[0::2].mean(), hop[1::2].mean(), comp[:-1].mean()))
            t_w = tl[:, 10].mean() * ghz; t_s = tl[:, 11].mean() * ghz
            print('   layer29 wake -> tail wake %.0f, tail wake->send %.0f' % (t_w - wake[29], t_s - t_w))


if __name__ == '__main__':
    main()
