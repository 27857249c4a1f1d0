"""Debug aid: fused tcgen05 layer kernel vs the cuBLASLt path (same bf16 inputs) and vs the fp32 oracle."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import train_oracle as to                      # noqa: E402
from tacotron_wavenet_vocoder_korean_b200 import synth     # noqa: E402
from tacotron_wavenet_vocoder_korean_b200.wavenet.train import WaveNetTrainer   # noqa: E402
from tests.train_helpers import train_case, rel_err, cosine, well_conditioned   # noqa: E402

kw = synth.cfg2(2)
T = 3600
w, wav, mel, gc = train_case(kw, T)
if os.environ.get('WELL_CONDITIONED'):
    w = well_conditioned(w)
res = {}
for mode in ('unfused', 'fused'):
    if mode == 'unfused':
        os.environ['WNT_NO_FUSED'] = '1'
    else:
        os.environ.pop('WNT_NO_FUSED', None)
    tr = WaveNetTrainer(T, dtype='bf16', **kw)
    tr.load_state_dict(w)
    L = float(tr.loss_and_grads(wav, mel, gc).item())
    res[mode] = dict(L=L, raw=tr.debug_get('raw_output'), x1=tr.debug_get('x1'), x7=tr.debug_get('x7'), x29=tr.debug_get('x29'),
                     g=tr.state_dict('grads'), info=tr.info())
    print(mode, 'loss', L, 'fused_launches', res[mode]['info']['fused_launches'], 'gemms', res[mode]['info']['gemm_launches'], flush=True)
a, b = res['unfused'], res['fused']
for k in ('x1', 'x7', 'x29', 'raw'):
    print(k, 'max abs diff', float(np.abs(a[k] - b[k]).max()), 'rel', rel_err(b[k], a[k]), 'max |ref|', float(np.abs(a[k]).max()))
worst = min((cosine(a['g'][k], b['g'][k]), k) for k in a['g'] if np.linalg.norm(a['g'][k]) > 1e-4)
print('worst gradient cosine fused vs unfused', worst)
low = sorted((cosine(a['g'][k], b['g'][k]), k) for k in a['g'] if np.linalg.norm(a['g'][k]) > 1e-4)[:8]
for c, k in low:
    print('   %.5f %s' % (c, k))
import collections
bykind = collections.defaultdict(list)
for k in a['g']:
    if np.linalg.norm(a['g'][k]) > 1e-4:
        bykind[k.split('/')[-2] + '/' + k.split('/')[-1]].append(cosine(a['g'][k], b['g'][k]))
for kk, v in sorted(bykind.items()):
    print('   kind %-28s min cos %.5f over %d tensors' % (kk, min(v), len(v)))
Lo, go = to.TorchWaveNetTrain(w, **kw).loss_and_grads(wav, mel, gc)
print('oracle loss', Lo)
for mode in res:
    c = min((cosine(res[mode]['g'][k], go[k]), k) for k in go if np.linalg.norm(go[k]) > 1e-4)
    print(mode, 'worst cosine vs fp32 oracle', c)

import collections
for mode in res:
    bk = collections.defaultdict(list)
    for k in go:
        if np.linalg.norm(go[k]) > 1e-4:
            bk[k.split('/')[-2] + '/' + k.split('/')[-1]].append((cosine(res[mode]['g'][k], go[k]), rel_err(res[mode]['g'][k], go[k])))
    print(mode, 'vs fp32 oracle, per kind: min cosine / max rel err')
    for kk, v in sorted(bk.items()):
        print('   %-24s %.5f  %.4f' % (kk, min(x[0] for x in v), max(x[1] for x in v)))
