import sys, os, ctypes as C, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tacotron_wavenet_vocoder_korean_b200 import _lib, synth
from tacotron_wavenet_vocoder_korean_b200.wavenet import WaveNetModel
lib = _lib.lib()
lib.wn_debug_calib_note.restype = C.c_char_p
torch.zeros(1).cuda()
kw = synth.cfg2(8); t0 = time.time(); net = WaveNetModel(train_mode=False, **kw); net.load_state_dict(synth.make_weights(**kw)); print('finalize s', time.time() - t0)
print('note:', lib.wn_debug_calib_note(), 'die_aware', net.info()['die_aware'])
out = np.zeros(148 * 16, np.uint32)
n = lib.wn_debug_calib_raw(out.ctypes.data_as(C.c_void_p), 148)
lat = out.reshape(148, 16)
np.set_printoptions(linewidth=220)
print(lat[1:6]); print(lat[140:148])
t0 = time.time(); net2 = WaveNetModel(train_mode=False, **kw); net2.load_state_dict(synth.make_weights(**kw)); print('second finalize s', time.time() - t0, net2.info()['die_aware'])
