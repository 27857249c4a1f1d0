# coding: utf-8
"""Generates tests/golden/ref_*.npz by running THE REFERENCE'S OWN wavenet/model.py, wavenet/mixture.py and wavenet/ops.py
(imported from /root/reference, unmodified) on top of the numpy TensorFlow stand-in tests/golden/tf_numpy_shim.py.
Run in the build container only (the GPU box has no /root/reference):

    python tests/golden/make_reference_goldens.py

Cases (seeded inputs = tests/helpers.make_inputs, weights = synth.make_weights, i.e. exactly what the parity tests feed):
  ref_mol     tiny MoL model (scalar input, mel + speaker conditioning): per-step network output (the tensor handed to
              sample_from_discretized_mix_logistic) and the drawn sample under teacher forcing, for 64 steps;
              create_upsample output; calculate_receptive_field.
  ref_mulaw   tiny mu-law model with mel + speaker conditioning: per-step softmax probabilities (predict_proba_incremental).
  ref_cfg2    the benchmark configuration (BASELINE configs[1] layer sizes), 2 rows x 640 teacher-forced steps: dilation 512 reads
              non-zero delayed taps from t = 512 on, through the reference's own queue code (model.py:49-64,145).
  ref_train   add_loss (train mode) of the tiny training model: the scalar loss the reference graph evaluates, with and
              without L2, and mu_law_encode / mu_law_decode of an amplitude grid (wavenet/ops.py).
  ref_train_onehot  add_loss for scalar_input=False (mu-law one-hot input, softmax cross-entropy head), tiny models with
              and without conditioning.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get('REFERENCE_ROOT', '/root/reference')

sys.path.insert(0, HERE)
import tf_numpy_shim as tf      # noqa: E402

tf.install()
sys.path.insert(0, REF)
import wavenet as ref_wavenet   # noqa: E402  (the reference package: /root/reference/wavenet)
from wavenet import model as ref_model, mixture as ref_mixture   # noqa: E402

assert os.path.abspath(ref_wavenet.__file__).startswith(os.path.abspath(REF)), ref_wavenet.__file__
sys.path.insert(1, ROOT)
from tacotron_wavenet_vocoder_korean_b200 import synth    # noqa: E402
from tests.helpers import make_inputs                       # noqa: E402
from tests.train_helpers import train_case, at_cell_centres  # noqa: E402


def build(kw, train_mode):
    tf.reset()
    w = synth.make_weights(**kw)
    tf.set_initial_values(w)
    net = ref_wavenet.WaveNetModel(train_mode=train_mode, **kw)
    return net, w


def incremental_case(kw, T):
    net, w = build(kw, False)
    inp = make_inputs(kw, T)
    with tf.graph_pass():
        with tf.variable_scope('wavenet', reuse=tf.AUTO_REUSE):        # generate.py:154-155
            lc_up = np.array(net.create_upsample(inp['mel']))
    captured = []
    orig = ref_model.sample_from_discretized_mix_logistic

    def spy(y, *a, **k):
        captured.append(np.array(y))
        return orig(y, *a, **k)
    ref_model.sample_from_discretized_mix_logistic = spy
    outs = []
    try:
        for t in range(T):
            x = inp['forced_full'][:, t:t + 1]
            if kw['scalar_input']:
                u = inp['uniforms'][:, t]
                tf.push_uniforms(u[:, None, :-1], u[:, -1:])              # mixture.py:103 (N,1,nr_mix) then :110 (N,1)
            else:
                x = x.astype(np.int32)
            with tf.graph_pass():                                       # one sess.run(next_sample, feed_dict) of generate.py:209
                out = net.predict_proba_incremental(x, lc_up[:, t, :], inp['gc_ids'])
            outs.append(np.array(out))
    finally:
        ref_model.sample_from_discretized_mix_logistic = orig
    names = sorted(n for n in tf.S.created_order if 'queue' not in n)
    assert set(names) == set(w), (set(names) ^ set(w))                  # every name, nothing else
    queues = [n for n in tf.S.created_order if 'queue' in n]
    res = dict(lc_up=lc_up, outputs=np.stack(outs, 1), variable_names=np.array(names), queue_names=np.array(queues),
               queue_shapes=np.array([str(tuple(tf.S.variables[n].shape)) for n in queues]),
               receptive_field=np.int64(net.receptive_field))
    if captured:
        res['raw_output'] = np.stack([c[:, 0] for c in captured], 1)    # (N, T, out_channels)
    return res


def train_case_loss(kw, T, codec=True, snap=False):
    out = {}
    w, wav, mel, gc = train_case(kw, T)
    if snap:
        wav = at_cell_centres(wav, kw['quantization_channels'])
    for tag, l2 in (('loss', None), ('loss_l2', 0.01)):
        net, _ = build(kw, True)
        with tf.graph_pass():
            net.add_loss(input_batch=wav[:, :, None], local_condition=mel, global_condition_batch=gc, l2_regularization_strength=l2)
        out[tag] = np.float64(net.loss)
    if not codec:
        return out
    grid = np.linspace(-1.2, 1.2, 4001).astype(np.float32)
    enc = np.array(ref_wavenet.mu_law_encode(grid, 256))
    out.update(mu_grid=grid, mu_encoded=enc.astype(np.int32), mu_decoded=np.array(ref_wavenet.mu_law_decode(np.arange(256, dtype=np.int32), 256)).astype(np.float32))
    return out


def main():
    kw = synth.tiny_mol()
    np.savez_compressed(os.path.join(HERE, 'ref_mol.npz'), **incremental_case(kw, 64))
    kw = dict(synth.tiny_mulaw(), local_condition_channels=20, upsample_factor=[2, 3], global_condition_channels=8, global_condition_cardinality=3)
    np.savez_compressed(os.path.join(HERE, 'ref_mulaw.npz'), **incremental_case(kw, 48))
    np.savez_compressed(os.path.join(HERE, 'ref_train.npz'), **train_case_loss(synth.tiny_train(3), 96))
    np.savez_compressed(os.path.join(HERE, 'ref_train_cfg2.npz'), **train_case_loss(synth.cfg2(2), 3600, codec=False))   # BASELINE configs[3] layers
    # scalar_input=False: mu-law one-hot input + softmax cross-entropy (model.py:257-296), with and without conditioning
    np.savez_compressed(os.path.join(HERE, 'ref_train_onehot.npz'),
                        **{k + '_lc_gc': v for k, v in train_case_loss(dict(synth.tiny_train(3), scalar_input=False), 96, codec=False, snap=True).items()},
                        **train_case_loss(synth.tiny_mulaw(2), 96, codec=False, snap=True))
    # BASELINE configs[1] layer sizes (30 layers, R=D=128, S=512, MoL-10, 80-channel mel, 2 speakers), 2 rows x 640 steps
    g = incremental_case(synth.cfg2(2), 640)
    g['lc_up'] = g['lc_up'][:, :640]
    del g['variable_names']
    np.savez_compressed(os.path.join(HERE, 'ref_cfg2.npz'), **g)
    rf = [ref_wavenet.WaveNetModel.calculate_receptive_field(2, [1, 2, 4, 8, 16, 32, 64, 128, 256, 512] * 5, s, 32) for s in (False, True)]
    print('receptive fields (non-scalar, scalar, 50 layers):', rf)
    for f in ('ref_mol', 'ref_mulaw', 'ref_train', 'ref_cfg2'):
        g = np.load(os.path.join(HERE, f + '.npz'))
        print(f, {k: g[k].shape for k in g.files})


if __name__ == '__main__':
    main()
