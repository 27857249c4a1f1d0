"""Shared builders for the Tacotron parity tests: the same seeded inputs go to the numpy oracle
(oracle/taco_oracle.py) and to the CUDA path (through the C ABI)."""
import numpy as np

from tacotron_wavenet_vocoder_korean_b200 import synth


def make_batch(N, T_in, num_symbols=80, seed=0, min_len=4):
    """Random jamo ids in [2, num_symbols), EOS (1) at the end of each sentence, 0-padded to T_in."""
    rs = np.random.RandomState(seed)
    ids = rs.randint(2, num_symbols, (N, T_in)).astype(np.int32)
    lens = rs.randint(min(min_len, T_in), T_in + 1, (N,)).astype(np.int32)
    lens[0] = T_in
    for n in range(N):
        ids[n, lens[n] - 1] = 1
        ids[n, lens[n]:] = 0
    spk = (np.arange(N) % 2).astype(np.int32)
    return ids, lens, spk


class Bag(object):
    """hparams-like attribute bag over a dict."""
    def __init__(self, d):
        self.__dict__.update(d)


CASES = {
    # name: (hp overrides on synth.taco_tiny(), num_speakers, N, T_in, steps)
    'tiny_mon_norm': (dict(), 2, 3, 11, 12),
    'tiny_mon': (dict(attention_type='bah_mon'), 2, 2, 9, 8),
    'tiny_loc_sen': (dict(attention_type='loc_sen'), 2, 3, 13, 10),
    'tiny_single_speaker': (dict(), 1, 2, 10, 8),
    'tiny_two_tiles': (dict(), 2, 35, 9, 6),
    'tiny_post_dense': (dict(post_proj_sizes=[32, 20], post_rnn_size=24, enc_rnn_size=16), 2, 2, 8, 5),
}


# the reference's hparams.py:124-166 layer sizes (7.07 M parameters) on short inputs
FULL_CASES = {
    'full_mon_norm': (dict(), 2, 2, 14, 6, 21),
    'full_loc_sen': (dict(attention_type='loc_sen'), 2, 2, 12, 5, 22),
}


def case(name):
    if name in FULL_CASES:
        over, ns, N, T_in, steps, seed = FULL_CASES[name]
        hp = dict(synth.TACO_HP, **over)
        w = synth.make_taco_weights(hp, ns)
        ids, lens, spk = make_batch(N, T_in, hp['num_symbols'], seed=seed)
        return hp, ns, w, ids, lens, spk, steps
    over, ns, N, T_in, steps = CASES[name]
    hp = synth.taco_tiny(**over)
    w = synth.make_taco_weights(hp, ns, seed=4321)
    ids, lens, spk = make_batch(N, T_in, hp['num_symbols'], seed=len(name))
    return hp, ns, w, ids, lens, (spk if ns > 1 else None), steps
