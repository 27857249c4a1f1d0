"""Host-side logic that needs no GPU: weight initialisers."""
import numpy as np


def test_training_init_is_glorot_everywhere_and_seedable():
    """A fresh training run starts from tf.global_variables_initializer's distribution (train_vocoder.py:129-130): Glorot-uniform
    kernels including the conv2d_transpose upsamplers, zero biases -- not from the benchmark's hold-filter heuristic."""
    from tacotron_wavenet_vocoder_korean_b200 import synth
    kw = synth.tiny_mol(2)
    a = synth.make_weights(seed=1, init='train', **kw)
    b = synth.make_weights(seed=2, init='train', **kw)
    bench = synth.make_weights(**kw)
    up = [k for k in a if 'upsample' in k and k.endswith('kernel')]
    assert up
    for k in up:
        shp = a[k].shape
        rf = int(np.prod(shp[:-2]))
        lim = np.sqrt(6.0 / (rf * shp[-2] + rf * shp[-1]))
        assert np.abs(a[k]).max() <= lim + 1e-7 and a[k].min() < 0 < a[k].max()
        assert bench[k].min() >= 0.3                       # the benchmark heuristic is untouched
        assert not np.array_equal(a[k], b[k])
    assert all(not a[k].any() for k in a if k.endswith('/bias'))
