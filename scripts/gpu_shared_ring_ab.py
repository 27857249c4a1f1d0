"""A/B of the shared tagged dilation ring (WN_SHARED_RING=1) against four private ring copies per layer (=0) on the cluster path:
us per step at 8 / 12 / 16 / 24 / 32 rows in flight, same process, alternating.  usage: gpu_shared_ring_ab.py [steps]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402
from tacotron_wavenet_vocoder_korean_b200 import synth  # noqa: E402
from tacotron_wavenet_vocoder_korean_b200.wavenet import WaveNetModel  # noqa: E402

T = int(sys.argv[1]) if len(sys.argv) > 1 else 6000
for rows in (8, 12, 16, 24, 32):
    kw = synth.cfg2(rows)
    net = WaveNetModel(train_mode=False, **kw)
    net.load_state_dict(synth.make_weights(**kw))
    rs = np.random.RandomState(100)
    mel = torch.from_numpy(np.clip(rs.randn(rows, (T + 299) // 300, 80) * 1.5, -4, 4).astype(np.float32)).cuda()
    uni = torch.from_numpy(rs.uniform(1e-5, 1 - 1e-5, (rows, T, 11)).astype(np.float32)).cuda()
    x0 = torch.from_numpy((2 * rs.rand(rows, 1) - 1).astype(np.float32)).cuda()
    gc = [i * 2 // rows for i in range(rows)]
    res = {}
    ref = None
    for flag in ('0', '1', '0', '1'):
        os.environ['WN_SHARED_RING'] = flag
        net.generate(T, x0, uni, mel=mel, gc_ids=gc)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = net.generate(T, x0, uni, mel=mel, gc_ids=gc, sync=False)
        e1.record()
        torch.cuda.synchronize()
        net.sync_check()
        ref = out if ref is None else ref
        assert torch.equal(out, ref), "shared / private ring outputs differ at %d rows" % rows
        res.setdefault(flag, []).append(1e3 * e0.elapsed_time(e1) / T)
    print("rows %2d x %d steps: private rings %s us/step, shared ring %s us/step -> %.0f k vs %.0f k samples/s" % (
        rows, T, " / ".join("%.2f" % v for v in res['0']), " / ".join("%.2f" % v for v in res['1']),
        rows / min(res['0']) * 1e3, rows / min(res['1']) * 1e3))
    del net
    torch.cuda.empty_cache()
