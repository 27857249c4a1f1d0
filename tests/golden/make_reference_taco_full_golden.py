# coding: utf-8
"""Generates tests/golden/ref_taco_full_<attention>.npz by running THE REFERENCE'S OWN tacotron package -- tacotron/tacotron.py
(`Tacotron.initialize`, inference mode), tacotron/rnn_wrappers.py (its AttentionWrapper with the manual-alignment override,
DecoderPrenetWrapper, ConcatOutputAndAttentionWrapper, LocationSensitiveAttention), tacotron/helpers.py (TacoTestHelper) and
tacotron/modules.py, imported unmodified from /root/reference -- on the numpy stand-ins tf_numpy_shim.py / tf_contrib_shim.py.

Everything the reference's code decides is exercised for real: the zeroed <PAD> embedding row, the deepvoice speaker
states, what is concatenated where, the order of the cell stack, the go frame and the feedback of the last of r frames,
the stop rule, the alignment history layout, the post net and the final projection.  tf.contrib's cells, attention
mechanisms and decode loop are restated in tf_contrib_shim.py (third-party code).  The decoder variable names that TF's
wrappers would generate are mapped onto this repository's short names by NAME_MAP below.

    python tests/golden/make_reference_taco_full_golden.py        (build container only)
"""
import os
import re
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get('REFERENCE_ROOT', '/root/reference')
sys.path.insert(0, HERE)
import tf_contrib_shim as contrib      # noqa: E402

tf = contrib.install()
sys.path.insert(1, ROOT)
from tacotron_wavenet_vocoder_korean_b200 import synth                      # noqa: E402
from tacotron_wavenet_vocoder_korean_b200.text.symbols import symbols as my_symbol_table   # noqa: E402

# the reference's `utils` and `text` packages pull in librosa / jamo at import: tacotron.py only needs `log` and `symbols`
u = types.ModuleType('utils')
u.infolog = types.ModuleType('utils.infolog')
u.infolog.log = print
sys.modules.update({'utils': u, 'utils.infolog': u.infolog})
txt = types.ModuleType('text')
txt.symbols = types.ModuleType('text.symbols')
txt.symbols.symbols = list(my_symbol_table)
sys.modules.update({'text': txt, 'text.symbols': txt.symbols})
sys.path.insert(0, REF)
import tacotron as ref_tacotron        # noqa: E402  (/root/reference/tacotron)

assert os.path.abspath(ref_tacotron.__file__).startswith(os.path.abspath(REF))

# TF-generated decoder scopes -> the short names of synth.taco_weight_shapes ("names assigned by tf.contrib
# wrappers -- unpinned")
P = 'model/inference/decoder/output_projection_wrapper/'
C0 = P + 'multi_rnn_cell/cell_0/output_projection_wrapper/'
AW = C0 + 'concat_output_and_attention_wrapper/decoder_prenet_wrapper/'
NAME_MAP = [
    (re.escape(AW) + r'decoder_prenet/(.*)', r'model/inference/decoder/decoder_prenet/\1'),
    (re.escape(AW) + r'attention_wrapper/gru_cell/(.*)', r'model/inference/decoder/attention_cell/gru_cell/\1'),
    (re.escape(AW) + r'attention_wrapper/(?:bahdanau_monotonic_attention|bahdanau_attention)/(.*)', r'model/inference/decoder/attention/\1'),
    (re.escape(AW) + r'attention_wrapper/Location_Sensitive_Attention/(.*)', r'model/inference/decoder/attention/\1'),
    (re.escape(C0) + r'(kernel|bias)', r'model/inference/decoder/concat_projection/\1'),
    (re.escape(P) + r'multi_rnn_cell/cell_(\d)/gru_cell/(.*)', r'model/inference/decoder/cell_\1/gru_cell/\2'),
    (re.escape(P) + r'(kernel|bias)', r'model/inference/decoder/output_projection/\1'),
]


def resolver(full):
    for pat, rep in NAME_MAP:
        if re.fullmatch(pat, full):
            return re.sub(pat, rep, full)
    return full


class HP(object):
    def __init__(self, d):
        self.__dict__.update(d)


def run(case_name, manual=False):
    from tests.taco_helpers import case
    hp, ns, w, ids, lens, spk, steps = case(case_name)
    contrib.FEED.clear()
    tag = case_name
    extra = {}
    if manual:      # is_manual_attention / manual_alignments of the feed_dict (synthesizer.py:138-150, rnn_wrappers.py:374)
        rs = np.random.RandomState(17)
        man = rs.rand(len(lens), steps, ids.shape[1]).astype(np.float32)
        man /= man.sum(-1, keepdims=True)
        contrib.FEED.update(is_manual_attention=True, manual_alignments=man)
        extra['manual_alignments'] = man
        tag += '_manual'
    tf.reset()
    tf.set_initial_values(w)
    tf.S.resolver = resolver
    # with one speaker the reference only works for model_type != 'deepvoice': tacotron.py:185 iterates decoder_rnn_init_states = None
    hpo = HP(dict(hp, max_iters=steps, dropout_prob=0.5, model_type=hp['model_type'] if ns > 1 else 'single'))
    model = ref_tacotron.create_model(hpo)
    with tf.graph_pass():
        with tf.variable_scope('model'):                                   # synthesizer.py:53
            model.initialize(np.asarray(ids), np.asarray(lens), ns, None if spk is None else np.asarray(spk), rnn_decoder_test_mode=True)   # synthesizer.py:54-56
    keep_lin = 64 if case_name.startswith('full') else None      # full size: the first 64 of the 1025 linear bins keep the fixture small
    out = dict(extra, ids=np.asarray(ids), lengths=np.asarray(lens), speaker_ids=np.asarray(spk if spk is not None else []), steps=np.int64(steps),
               mel_outputs=np.array(model.mel_outputs), linear_outputs=np.array(model.linear_outputs)[..., :keep_lin], alignments=np.array(model.alignments),
               tf_names=np.array(sorted(tf.S.resolved)), mapped_names=np.array([tf.S.resolved[k] for k in sorted(tf.S.resolved)]))
    unused = sorted(set(w) - set(tf.S.resolved.values()))
    assert not unused, ('weights the reference graph never asked for', unused)
    np.savez_compressed(os.path.join(HERE, 'ref_taco_full_%s.npz' % tag), **out)
    print(tag, {k: np.shape(v) for k, v in out.items() if k not in ('tf_names', 'mapped_names')}, len(tf.S.resolved), 'variables')


if __name__ == '__main__':
    for name in (sys.argv[1:] or ['tiny_mon_norm', 'tiny_mon', 'tiny_loc_sen', 'tiny_single_speaker', 'tiny_post_dense', 'full_mon_norm', 'full_loc_sen']):
        run(name)
    if not sys.argv[1:]:
        run('tiny_mon_norm', manual=True)
