"""Generates the committed golden fixtures (tests/golden/*.npz) from the CPU oracle.

The reference cannot run here (no TensorFlow 1.x), so these are outputs of the *restatement*
(oracle/wn_oracle.c) on seeded inputs -- they pin the oracle against drift and give the GPU tests
fixed vectors to match; they are NOT outputs of the reference itself (parity unpinned).

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import oracle                                                            # noqa: E402
from oracle import np_oracle                                             # noqa: E402
from tests.helpers import make_inputs, oracle_model, plan_from_dict     # noqa: E402
from tacotron_wavenet_vocoder_korean_b200 import synth                   # noqa: E402
from tacotron_wavenet_vocoder_korean_b200.wavenet.model import plan_config   # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))

CASES = {
    # name: (config factory, kwargs, T, teacher-forced steps kept with logits)
    'tiny_mol': (synth.tiny_mol, {}, 120),
    'tiny_mulaw': (synth.tiny_mulaw, {}, 200),
    'cfg1': (synth.cfg1, {}, 600),
    'cfg2_n2': (synth.cfg2, {'batch_size': 2}, 96),
}


def run_case(name):
    fac, fkw, T = CASES[name]
    kw = fac(**fkw)
    w = synth.make_weights(**kw)
    om = oracle_model(kw, w)
    inp = make_inputs(kw, T)
    plan, _ = plan_config(**kw)
    lc = om.upsample(inp['mel']) if 'mel' in inp else None
    out = {'plan': np.array([plan[k] for k, _ in oracle.OrcPlan._fields_], np.int32), 'T': np.int32(T)}
    for tag, pl in (('kernel', plan_from_dict(plan)), ('natural', oracle.OrcPlan.natural())):
        s, lg = om.generate(T, inp['x0'], inp['uniforms'], lc_up=lc, gc_ids=inp['gc_ids'], plan=pl, want_logits=True)
        out['samples_' + tag] = s
        out['logits_tail_' + tag] = lg[:, -4:, :]
        s2, lg2 = om.generate(min(T, 48), inp['forced_full'][:, :min(T, 48)], inp['uniforms'][:, :min(T, 48)],
                              lc_up=lc, gc_ids=inp['gc_ids'], plan=pl, want_logits=True)
        out['tf_logits_' + tag] = lg2
    if lc is not None:
        out['lc_checksum'] = np.array([np.float64(lc.astype(np.float64).sum()), np.float64(np.abs(lc).astype(np.float64).sum())])
        out['lc_head'] = lc[:, :7, :]
    np.savez_compressed(os.path.join(HERE, name + '.npz'), **out)
    print(name, {k: getattr(v, 'shape', None) for k, v in out.items()})


def codec_tables():
    grid = np.linspace(-1.2, 1.2, 4801).astype(np.float32)
    enc = oracle.mu_law_encode(grid, 256)
    dec_q = oracle.mu_law_decode(np.arange(256, dtype=np.float32), 256, True)
    dec_c = oracle.mu_law_decode(np.linspace(-1, 1, 513).astype(np.float32), 256, False)
    rf = np.array([
        np_oracle.calculate_receptive_field(2, [1, 2, 4, 8, 16, 32, 64, 128, 256, 512] * 5, False, 32),   # 5117, generate.py:192
        np_oracle.calculate_receptive_field(2, [1, 2, 4, 8, 16, 32, 64, 128, 256, 512] * 5, True, 32),    # 5147
        np_oracle.calculate_receptive_field(2, [1, 2, 4, 8, 16, 32, 64, 128, 256, 512], False, 32),       # 1025
        np_oracle.calculate_receptive_field(2, [1, 2, 4, 8, 16, 32, 64, 128, 256, 512] * 3, True, 32),    # 3101
    ], np.int32)
    np.savez_compressed(os.path.join(HERE, 'codec.npz'), grid=grid, enc=enc, dec_q=dec_q, dec_c=dec_c, rf=rf)
    print('codec', enc[:3], enc[-3:], rf)


if __name__ == '__main__':
    for n in CASES:
        run_case(n)
    codec_tables()
