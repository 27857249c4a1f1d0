"""ncu target: one warm-up and one profiled `taco_synthesize` of cfg-3 (32 sentences, 200 decoder steps).
ncu ... --launch-skip <launches of the warm-up> python scripts/ncu_taco_target.py"""
import os
import sys

import numpy as np

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
steps = int(os.environ.get('TACO_STEPS', '200'))
for _ in range(2):
    m.initialize(ids, lens, 2, spk, rnn_decoder_test_mode=True, n_steps=steps)
    torch.cuda.synchronize()
print("launches per call:", m.info()['kernel_launches'] // 2)
