"""Target for `ncu --replay-mode range`: one generation call of the benchmark shape between cudaProfilerStart/Stop.
The cluster path runs two CONCURRENT kernels (layer clusters + tail) that talk through L2 mailboxes; ncu's default
kernel replay serialises launches and would deadlock them (the in-kernel watchdog then aborts), so hardware counters
are collected over the whole range instead.  usage: ncu_range_target.py [rows] [steps] [fast]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tacotron_wavenet_vocoder_korean_b200 import synth
from tacotron_wavenet_vocoder_korean_b200.wavenet import WaveNetModel

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 8
T = int(sys.argv[2]) if len(sys.argv) > 2 else 3000
fast = len(sys.argv) > 3 and sys.argv[3] == 'fast'
kw = synth.cfg2(rows)
net = WaveNetModel(train_mode=False, fast_act=fast, **kw)
net.load_state_dict(synth.make_weights(**kw))
rs = np.random.RandomState(100)
mel = torch.from_numpy(np.clip(rs.randn(rows, (T + 299) // 300, 80) * 1.5, -4, 4).astype(np.float32)).cuda()
uni = torch.from_numpy(rs.uniform(1e-5, 1 - 1e-5, (rows, T, 11)).astype(np.float32)).cuda()
x0 = torch.from_numpy((2 * rs.rand(rows, 1) - 1).astype(np.float32)).cuda()
gc = [i * 2 // rows for i in range(rows)]
net.generate(T, x0, uni, mel=mel, gc_ids=gc)           # warm-up outside the range
torch.cuda.synchronize()
torch.cuda.profiler.start()
out = net.generate(T, x0, uni, mel=mel, gc_ids=gc, sync=False)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
net.sync_check()
print('range done: rows %d steps %d, sample std %.4f, info %s' % (rows, T, float(out.std()), net.info()))
