# coding: utf-8
"""Generates tests/golden/ref_optimizer.json by running THE REFERENCE'S OWN WaveNetModel.add_optimizer (wavenet/model.py:314-346) and
the EMA construction (:30) with recording stand-ins for tf.train.*: which schedule, optimizer, clipping and averaging it asks
TensorFlow for, with which arguments (the hparams.py values of the reference).  The arithmetic of Adam / exponential_decay /
clip_by_global_norm / ExponentialMovingAverage is TensorFlow's (third party; pinned against torch.optim and closed forms in
tests/test_train_oracle.py); this pins the PLUMBING: non-staircase decay of wavenet_learning_rate, Adam with TF's default betas,
clip norm 1.0 behind wavenet_clip_gradients, EMA over every trainable variable with the decay of model.py:30.

    python tests/golden/make_reference_optimizer_golden.py        (build container only)
"""
import json
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get('REFERENCE_ROOT', '/root/reference')
sys.path.insert(0, HERE)
import make_reference_audio_golden as mra      # noqa: E402
import tf_numpy_shim as tf                     # noqa: E402


def main():
    mra.install_stubs()
    sys.path.insert(0, REF)
    sys.path.insert(1, ROOT)
    from tacotron_wavenet_vocoder_korean_b200 import synth
    import wavenet as ref_wavenet
    hp = mra.load(os.path.join(REF, 'hparams.py'), 'ref_hparams').hparams
    rec = {'calls': []}

    class Adam(object):
        def __init__(self, *a, **k):
            rec['adam_args'] = [str(x) for x in a]
            rec['adam_kwargs'] = {kk: vv for kk, vv in k.items()}

        def compute_gradients(self, loss):
            rec['calls'].append('compute_gradients')
            return [('grad:' + n, tf.S.variables[n]) for n in tf.S.trainable]

        def apply_gradients(self, gv, global_step=None):
            rec['calls'].append('apply_gradients')
            rec['applied_clipped'] = all(str(g).startswith('clipped:') for g, _ in gv)
            rec['apply_has_global_step'] = global_step is not None
            return 'adam_optimize'

    class EMA(object):
        def __init__(self, decay, **k):
            rec['ema_decay'] = decay
            rec['ema_kwargs'] = k

        def apply(self, variables):
            rec['calls'].append('ema.apply')
            rec['ema_over_all_trainables'] = len(variables) == len(tf.S.trainable)
            return 'optimize'

    def exponential_decay(lr, global_step, decay_steps, decay_rate, *a, **k):
        rec['exponential_decay'] = dict(learning_rate=lr, decay_steps=decay_steps, decay_rate=decay_rate, extra_args=list(a), extra_kwargs=k)
        return 'lr'

    def clip_by_global_norm(grads, norm):
        rec['clip_norm'] = norm
        return ['clipped:' + str(g) for g in grads], None

    tf.train.AdamOptimizer = Adam
    tf.train.ExponentialMovingAverage = EMA
    tf.train.exponential_decay = exponential_decay
    tf.clip_by_global_norm = clip_by_global_norm
    tf.get_collection = lambda *_: []
    tf.GraphKeys = types.SimpleNamespace(UPDATE_OPS='update_ops')
    out = {}
    for clip in (True, False):
        rec.clear()
        rec['calls'] = []
        kw = synth.tiny_train(2)
        tf.reset()
        tf.set_initial_values(synth.make_weights(**kw))
        net = ref_wavenet.WaveNetModel(train_mode=True, **kw)
        from tests.train_helpers import train_case
        w, wav, mel, gc = train_case(kw, 72)
        with tf.graph_pass():
            net.add_loss(input_batch=wav[:, :, None], local_condition=mel, global_condition_batch=gc, l2_regularization_strength=None)
            hp.wavenet_clip_gradients = clip
            net.add_optimizer(hp, 'global_step')
        assert net.optimize == 'optimize' and net.learning_rate == 'lr'
        out['clip_%s' % clip] = dict(rec)
    out['hparams'] = {k: getattr(hp, k) for k in ('wavenet_learning_rate', 'wavenet_decay_steps', 'wavenet_decay_rate', 'wavenet_clip_gradients',
                                                  'l2_regularization_strength', 'wavenet_batch_size', 'sample_size')
                      if hasattr(hp, k)}
    json.dump(out, open(os.path.join(HERE, 'ref_optimizer.json'), 'w'), indent=1, sort_keys=True)
    print(json.dumps(out, indent=1, sort_keys=True))


if __name__ == '__main__':
    main()
