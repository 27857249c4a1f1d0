"""Generates tests/golden/taco_*.npz: outputs of the numpy Tacotron oracle (oracle/taco_oracle.py, fp32) on the
seeded cases of tests/taco_helpers.py.  The reference itself cannot run here (no TensorFlow), so these fixtures
pin the ORACLE against accidental change and give the GPU tests a checker that does not need the oracle's code
path to stay bit-stable.  Run from the repo root: python tests/golden/make_golden_taco.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.taco_oracle import TacotronOracle  # noqa: E402
from tests.taco_helpers import case  # noqa: E402

for name in ('tiny_mon_norm', 'tiny_loc_sen'):
    hp, ns, w, ids, lens, spk, steps = case(name)
    mel, lin, al = TacotronOracle(hp, w, ns).synthesize(ids, lens, spk, max_iters=steps)
    np.savez_compressed(os.path.join(ROOT, 'tests', 'golden', 'taco_%s.npz' % name), ids=ids, lens=lens,
                        spk=spk if spk is not None else np.zeros(0, np.int32), mel=mel, linear=lin, alignments=al)
    print(name, mel.shape, lin.shape, al.shape)
