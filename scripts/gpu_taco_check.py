"""Stage-by-stage parity report of the Tacotron CUDA path against the numpy oracle (prints max |err| per stage
instead of stopping at the first failure).  Run on the GPU box: python scripts/gpu_taco_check.py"""
import os
import sys
import time
import traceback

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from oracle.taco_oracle import TacotronOracle  # noqa: E402
from tacotron_wavenet_vocoder_korean_b200 import synth  # noqa: E402
from tacotron_wavenet_vocoder_korean_b200.tacotron import Tacotron  # noqa: E402
from tests.taco_helpers import Bag, CASES, case, make_batch  # noqa: E402

PAIRS = [('enc_prenet', 'enc_prenet'), ('enc_bank', 'encoder_cbhg/bank'), ('enc_highway_in', 'encoder_cbhg/highway_in'),
         ('enc_rnn_in', 'encoder_cbhg/rnn_in'), ('encoder_out', 'encoder_out'), ('post_bank', 'post_cbhg/bank'),
         ('post_highway_in', 'post_cbhg/highway_in'), ('post_rnn_in', 'post_cbhg/rnn_in'), ('post_out', 'post_out')]


def report(name, hp, ns, w, ids, lens, spk, steps):
    print("== %s: N=%d T_in=%d steps=%d att=%s" % (name, ids.shape[0], ids.shape[1], steps, hp['attention_type']), flush=True)
    taps = {}
    t0 = time.time()
    mel, lin, al = TacotronOracle(hp, w, ns).synthesize(ids, lens, spk, max_iters=steps, taps=taps)
    t_or = time.time() - t0
    m = Tacotron(Bag(hp))
    m.load_state_dict(w)
    m.initialize(ids, lens, ns, spk, rnn_decoder_test_mode=True, n_steps=steps)
    torch.cuda.synchronize()
    for cname, oname in PAIRS:
        try:
            ref = taps[oname]
            got = m.debug_tensor(cname, ref.shape)
            print("   %-16s max|err| %.3g   (max|ref| %.3g)" % (cname, np.abs(got - ref).max(), np.abs(ref).max()))
        except Exception as e:  # noqa: BLE001
            print("   %-16s FAILED: %s" % (cname, e))
    gm, ga, gl = m.mel_outputs.cpu().numpy(), m.alignments.cpu().numpy(), m.linear_outputs.cpu().numpy()
    r = hp['reduction_factor']
    per_step = np.abs(gm - mel).reshape(gm.shape[0], steps, -1).max(axis=(0, 2))
    print("   mel max|err| %.3g  alignments %.3g  linear %.3g   first steps: %s" %
          (np.abs(gm - mel).max(), np.abs(ga - al).max(), np.abs(gl - lin).max(), np.array2string(per_step[:6], precision=2)))
    print("   oracle %.2f s; info %s" % (t_or, m.info()), flush=True)


if __name__ == "__main__":
    for name in sorted(CASES):
        try:
            report(name, *case(name))
        except Exception:  # noqa: BLE001
            traceback.print_exc()
    try:
        hp = dict(synth.TACO_HP)
        w = synth.make_taco_weights(hp, 2)
        ids, lens, spk = make_batch(4, 40, seed=7)
        report('full_size', hp, 2, w, ids, lens, spk, 30)
    except Exception:  # noqa: BLE001
        traceback.print_exc()
