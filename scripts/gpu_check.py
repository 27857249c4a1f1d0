"""First-contact GPU check: kernel vs oracle on a ladder of configurations (debug aid)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import oracle
from tests.helpers import make_inputs, oracle_model, plan_from_dict
from tacotron_wavenet_vocoder_korean_b200 import synth
from tacotron_wavenet_vocoder_korean_b200.wavenet import WaveNetModel


def run(name, kw, T, force_M=0, force_Mt=0, teacher=False, check=True):
    w = synth.make_weights(**kw)
    net = WaveNetModel(train_mode=False, force_M=force_M, force_Mt=force_Mt, **kw)
    net.load_state_dict(w)
    plan = net.plan(); info = net.info()
    inp = make_inputs(kw, T)
    lc = None
    om = oracle_model(kw, w) if check else None
    if 'mel' in inp:
        lc_t = net.create_upsample(inp['mel'])
        torch.cuda.synchronize()
        lc = lc_t
        if check:
            lc_o = om.upsample(inp['mel'])
            print('  upsample bit-exact:', np.array_equal(lc_o, lc_t.cpu().numpy()), lc_o.shape)
    forced = inp['forced_full'] if teacher else inp['x0']
    t0 = time.time()
    s, lg = net.generate(T, forced, inp['uniforms'], lc_up=lc, gc_ids=inp['gc_ids'], want_logits=True)
    torch.cuda.synchronize()
    dt = time.time() - t0
    s = s.cpu().numpy(); lg = lg.cpu().numpy()
    msg = '%s: grid=%d M=%d Mt=%d T=%d N=%d  %.3fs  %.0f samples/s' % (name, info['grid'], info['M'], info['Mt'], T, kw['batch_size'], dt, T * kw['batch_size'] / dt)
    if check:
        so, lo = om.generate(T, forced, inp['uniforms'], lc_up=lc.cpu().numpy() if lc is not None else None,
                             gc_ids=inp['gc_ids'], plan=plan_from_dict(plan), want_logits=True)
        eq_s = np.array_equal(so, s); eq_l = np.array_equal(lo, lg)
        msg += '  samples bit-exact=%s logits bit-exact=%s maxdiff=%.3g' % (eq_s, eq_l, np.abs(lo - lg).max())
        if not eq_l:
            bad = np.argwhere(lo != lg)
            print('   first mismatch (b,t,o):', bad[0], lo[tuple(bad[0])], lg[tuple(bad[0])], 'n_bad', len(bad))
    print(msg, flush=True)


if __name__ == '__main__':
    print(torch.cuda.get_device_name(0))
    run('tiny_mol teacher', synth.tiny_mol(), 50, teacher=True)
    run('tiny_mol free', synth.tiny_mol(), 200)
    run('tiny_mol M2 Mt2', synth.tiny_mol(), 200, force_M=2, force_Mt=2)
    run('tiny_mulaw teacher', synth.tiny_mulaw(), 50, teacher=True)
    run('tiny_mulaw free', synth.tiny_mulaw(), 300)
    run('tiny_mulaw M4 Mt4', synth.tiny_mulaw(), 300, force_M=4, force_Mt=4)
    run('cfg1', synth.cfg1(), 2000)
    run('cfg2 N=1', synth.cfg2(1), 600)
    run('cfg2 N=8 short', synth.cfg2(8), 300)
    run('cfg2 N=8 perf', synth.cfg2(8), 6000, check=False)
    run('cfg2 N=1 perf', synth.cfg2(1), 6000, check=False)
    run('cfg1 perf', synth.cfg1(), 8000, check=False)
    run('hparams default', synth.cfg_hparams_default(2), 600)
