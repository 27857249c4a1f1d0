"""In-kernel phase profile of the persistent Tacotron decoder (TACO_PROFILE=1): cycles of thread 0 per CTA spent
staging inputs, in the dot-product loop, in dense phases overall, in the two attention phases and in grid barriers."""
import os
import sys

import numpy as np

os.environ['TACO_PROFILE'] = '1'
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch  # noqa: E402
from bench_taco import make_texts  # noqa: E402
from tacotron_wavenet_vocoder_korean_b200 import synth  # noqa: E402
from tacotron_wavenet_vocoder_korean_b200.tacotron import Tacotron  # noqa: E402
from tacotron_wavenet_vocoder_korean_b200.text import text_to_sequence, prepare_inputs  # noqa: E402
from tests.taco_helpers import Bag  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 32
steps = 200
hp = dict(synth.TACO_HP)
w = synth.make_taco_weights(hp, 2)
ids = prepare_inputs([text_to_sequence(t) for t in make_texts(n)])
lens = np.array([int(np.argmax(s == 1)) + 1 for s in ids], np.int32)
spk = (np.arange(n) % 2).astype(np.int32)
m = Tacotron(Bag(hp))
m.load_state_dict(w)
for _ in range(3):
    m.initialize(ids, lens, 2, spk, rnn_decoder_test_mode=True, n_steps=steps, want_linear=False)
    torch.cuda.synchronize()
G = m.info()['dec_grid']
prof = m.debug_tensor('dec_prof', (G, 16)).view(np.int64).reshape(G, 8).astype(np.float64)
names = ['stage', 'dot', 'dense_total', 'att_score', 'att_ctx', 'barrier', 'all', '-']
per_step = prof / steps
print("cycles per decoder step (mean / min / max over %d CTAs), N=%d:" % (G, n))
for i, nm in enumerate(names[:7]):
    print("  %-12s %9.0f %9.0f %9.0f" % (nm, per_step[:, i].mean(), per_step[:, i].min(), per_step[:, i].max()))
