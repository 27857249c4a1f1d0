import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tacotron_wavenet_vocoder_korean_b200 import _lib
lib = _lib.lib()
torch.zeros(1).cuda()
G, NG = 148, 8
part = np.array([1, 2, 3, 5, 8, 13, 21, 34, 55, 89], np.int32)
out = np.zeros(len(part) * NG * NG, np.int64); smids = np.zeros(G, np.uint32)
rc = lib.wn_debug_pingpong_grid(G, part.ctypes.data_as(C.c_void_p), len(part), NG, 100, out.ctypes.data_as(C.c_void_p), smids.ctypes.data_as(C.c_void_p))
print('rc', rc, 'cta0 smid', smids[0])
m = out.reshape(len(part), NG, NG)
np.set_printoptions(linewidth=200)
for i, k in enumerate(part):
    print('partner cta %d smid %d: rows X (partner inbox grain), cols Y (cta0 inbox grain)' % (k, smids[k]))
    print(m[i])
