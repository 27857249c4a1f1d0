# coding: utf-8
"""text/__init__.py:38-109 of the reference for the 'korean_cleaners' pipeline: text -> int32 ids of the
80-symbol jamo table with EOS (id 1) appended; symbols outside the table, PAD and EOS inside the text are dropped
(`_should_keep_symbol`, text/__init__.py:122-123)."""
import numpy as np

from .korean import tokenize, jamo_to_korean
from .symbols import symbols, PAD, EOS

_symbol_to_id = {s: i for i, s in enumerate(symbols)}
_id_to_symbol = {i: s for i, s in enumerate(symbols)}


def text_to_sequence(text, as_token=False):
    seq = [_symbol_to_id[s] for s in tokenize(text) if s in _symbol_to_id and s != PAD and s != EOS]
    seq.append(_symbol_to_id[EOS])
    if as_token:
        return sequence_to_text(seq, combine_jamo=True)
    return np.array(seq, dtype=np.int32)


def sequence_to_text(sequence, skip_eos_and_pad=False, combine_jamo=False):
    out = ''
    for i in sequence:
        s = _id_to_symbol.get(int(i))
        if s is not None and not (skip_eos_and_pad and s in (EOS, PAD)):
            out += s
    return jamo_to_korean(out) if combine_jamo else out


def prepare_inputs(sequences):
    """datasets/datafeeder_tacotron.py:288-290 `_prepare_inputs`: right-pad with 0 to the longest sequence."""
    T = max(len(s) for s in sequences)
    return np.stack([np.pad(np.asarray(s, np.int32), (0, T - len(s)), mode='constant') for s in sequences])
