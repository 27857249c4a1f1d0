"""TACO_TC_DEBUG=1: clock64 timeline of CTA (0,0,0) of every tensor-core conv / dense launch of one cfg-3 Tacotron call."""
import os
import sys

import numpy as np

os.environ['TACO_TC_DEBUG'] = '1'
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch  # noqa: E402
from bench_taco import make_texts  # noqa: E402
from tacotron_wavenet_vocoder_korean_b200 import synth  # noqa: E402
from tacotron_wavenet_vocoder_korean_b200.tacotron import Tacotron  # noqa: E402
from tacotron_wavenet_vocoder_korean_b200.text import text_to_sequence, prepare_inputs  # noqa: E402
from tests.taco_helpers import Bag  # noqa: E402

hp = dict(synth.TACO_HP)
w = synth.make_taco_weights(hp, 2)
ids = prepare_inputs([text_to_sequence(t) for t in make_texts(32)])
lens = np.array([int(np.argmax(s == 1)) + 1 for s in ids], np.int32)
spk = (np.arange(32) % 2).astype(np.int32)
m = Tacotron(Bag(hp))
m.load_state_dict(w)
for _ in range(3):
    m.initialize(ids, lens, 2, spk, rnn_decoder_test_mode=True, n_steps=20)
    torch.cuda.synchronize()
t = m.debug_tensor('tc_dbg', (32 * 16 * 2,)).view(np.int64).reshape(32, 16)
names = ['k-tiles', 'prologue', 'tma0 issued', 'tma1 issued', 'mma0 full', 'mma1 full', 'last full', 'accum ready', 'chunk0 done', 'chunk1 done', 'all done']
print('launch ' + ' '.join('%12s' % n for n in names))
for i, r in enumerate(t):
    if r[0]:
        print('%6d ' % i + ' '.join('%12d' % v for v in r[:11]))
