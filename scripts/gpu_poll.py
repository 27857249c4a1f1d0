import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tacotron_wavenet_vocoder_korean_b200 import _lib
lib = _lib.lib(); lib.wn_debug_pollbench.restype = C.c_longlong
torch.zeros(1).cuda()
for ctas in (1, 120):
    for warps, lanes, K in [(1, 1, 1), (1, 1, 4), (1, 32, 1), (1, 32, 4), (4, 32, 1), (4, 32, 4), (4, 32, 2), (8, 32, 4), (2, 32, 4), (4, 8, 4), (4, 16, 4)]:
        print('ctas %3d warps %d lanes %2d K %d : %5d cycles/round' % (ctas, warps, lanes, K, lib.wn_debug_pollbench(ctas, 2000, warps, lanes, K)), flush=True)
