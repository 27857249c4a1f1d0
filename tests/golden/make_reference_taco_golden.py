# coding: utf-8
"""Generates tests/golden/ref_taco_modules.npz by running THE REFERENCE'S OWN tacotron/modules.py (prenet, cbhg, highwaynet,
conv1d + batch norm; imported unmodified from /root/reference) on the numpy TensorFlow stand-in tests/golden/tf_numpy_shim.py:
the encoder CBHG (with the deepvoice speaker vectors, modules.py:47-51,66-69) and the post CBHG of the tiny Tacotron test
model, on seeded inputs with ragged lengths.  GRUCell / bidirectional_dynamic_rnn are tf.contrib code and are restated in the
stand-in; everything else (bank concat, max-pool, projections, residual + speaker vector, highway stack, variable names) is
decided by the reference's code.  The attention decoder (tacotron.py + rnn_wrappers.py) is built from tf.contrib.seq2seq
classes and is NOT covered.

    python tests/golden/make_reference_taco_golden.py        (build container only)
"""
import importlib.util
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get('REFERENCE_ROOT', '/root/reference')
sys.path.insert(0, HERE)
sys.path.insert(1, ROOT)
import tf_numpy_shim as tf            # noqa: E402

tf.install()
spec = importlib.util.spec_from_file_location('ref_taco_modules', os.path.join(REF, 'tacotron', 'modules.py'))
mod = importlib.util.module_from_spec(spec)
spec.loader.exec_module(mod)
from tacotron_wavenet_vocoder_korean_b200 import synth    # noqa: E402


def main():
    hp = synth.taco_tiny()
    w = synth.make_taco_weights(hp, 2)
    rs = np.random.RandomState(5)
    N, T = 3, 11
    lengths = np.array([11, 7, 4])
    x_emb = rs.randn(N, T, hp['embedding_size']).astype(np.float32)
    bh = rs.randn(N, hp['enc_prenet_sizes'][-1]).astype(np.float32) * 0.3
    init = rs.randn(N, 2 * hp['enc_rnn_size']).astype(np.float32) * 0.3
    tf.reset()
    tf.set_initial_values(w)
    out = dict(x_emb=x_emb, lengths=lengths, before_highway=bh, rnn_init=init)
    with tf.graph_pass():
        with tf.variable_scope('model'):
            with tf.variable_scope('inference'):
                pre = mod.prenet(x_emb, False, hp['enc_prenet_sizes'], 0.5, scope='prenet')       # tacotron.py:103
                enc = mod.cbhg(pre, lengths, False, hp['enc_bank_size'], hp['enc_bank_channel_size'], hp['enc_maxpool_width'],
                               hp['enc_highway_depth'], hp['enc_rnn_size'], hp['enc_proj_sizes'], hp['enc_proj_width'],
                               scope='encoder_cbhg', before_highway=bh, encoder_rnn_init_state=init)             # tacotron.py:105-112
                T2 = 13
                mel = rs.randn(N, T2, hp['num_mels']).astype(np.float32)
                post = mod.cbhg(mel, None, False, hp['post_bank_size'], hp['post_bank_channel_size'], hp['post_maxpool_width'],
                                hp['post_highway_depth'], hp['post_rnn_size'], hp['post_proj_sizes'], hp['post_proj_width'],
                                scope='post_cbhg')                                                               # tacotron.py:204-216
    out.update(prenet=np.array(pre), encoder_out=np.array(enc), mel=mel, post_out=np.array(post),
               variable_names=np.array(sorted(tf.S.created_order)))
    missing = [n for n in tf.S.created_order if n not in w]
    assert not missing, missing
    np.savez_compressed(os.path.join(HERE, 'ref_taco_modules.npz'), **out)
    print({k: np.shape(v) for k, v in out.items()})


if __name__ == '__main__':
    main()
