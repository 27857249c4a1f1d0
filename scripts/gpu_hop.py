import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tacotron_wavenet_vocoder_korean_b200 import _lib
lib = _lib.lib()
torch.cuda.init(); torch.zeros(1).cuda()
G = 148
for mode, name in [(0, 'relaxed.gpu'), (1, 'volatile'), (2, 'cg'), (3, 'relaxed x4 pollers')]:
    out = np.zeros(2 * G, np.int64)
    rc = lib.wn_debug_pingpong_all(G, 2000, mode, out.ctypes.data_as(C.c_void_p))
    rtt, smid = out[1:G], out[G:]
    print('%-20s rc=%d RTT cycles: min %d  p25 %d  median %d  p75 %d  max %d' % (name, rc, rtt.min(), np.percentile(rtt, 25), np.median(rtt), np.percentile(rtt, 75), rtt.max()))
    if mode == 0:
        print(' cta0 smid', smid[0])
        order = np.argsort(smid[1:]) + 1
        print(' RTT by partner smid:', [(int(smid[k]), int(out[k])) for k in order])
