"""Seeded synthetic WaveNet weights with the reference's TF variable names/shapes.

No checkpoint ships with the reference and TF's tensor-bundle format is not readable
here, so benchmarks and parity tests use random-initialised weights of the exact
architecture.  Initialisers follow the TF defaults the
reference relies on: Glorot-uniform kernels (tf.layers), Xavier-normal gc_embedding
(wavenet/model.py:194); biases get a small uniform perturbation instead of TF's zeros
so that bias handling is actually exercised, and the ten MoL log-scale biases are
shifted by -3 so synthetic samples are not saturated at +-1.
"""
import numpy as np


def weight_shapes(batch_size=None, dilations=(), filter_width=2, residual_channels=32, dilation_channels=32,
                  skip_channels=512, quantization_channels=256, out_channels=30, use_biases=False,
                  scalar_input=False, initial_filter_width=32, global_condition_channels=None,
                  global_condition_cardinality=None, local_condition_channels=80, upsample_factor=None,
                  **_ignored):
    """Ordered {tf_variable_name: shape} for the generation graph (model.py:41-165,194)."""
    R, D, S = residual_channels, dilation_channels, skip_channels
    fw = filter_width
    shapes = {}
    if global_condition_channels and global_condition_cardinality:
        shapes['wavenet/gc_embedding'] = (global_condition_cardinality, global_condition_channels)
    if local_condition_channels and upsample_factor:
        for i, f in enumerate(upsample_factor):
            shapes['wavenet/upsample%d/kernel' % i] = (f, fw, 1, 1)
    shapes['wavenet/conv1d/kernel'] = (initial_filter_width, 1, R) if scalar_input else (fw, quantization_channels, R)
    for l in range(len(dilations)):
        p = 'wavenet/dilated_stack/layer%d/dilation_layer/' % l
        shapes[p + 'conv_filter/kernel'] = (fw, R, D)
        shapes[p + 'conv_gate/kernel'] = (fw, R, D)
        if use_biases:
            shapes[p + 'conv_filter/bias'] = (D,)
            shapes[p + 'conv_gate/bias'] = (D,)
        if global_condition_channels:
            shapes[p + 'gc_filter/kernel'] = (1, global_condition_channels, D)
            shapes[p + 'gc_gate/kernel'] = (1, global_condition_channels, D)
        if local_condition_channels:
            shapes[p + 'lc_filter/kernel'] = (1, local_condition_channels, D)
            shapes[p + 'lc_gate/kernel'] = (1, local_condition_channels, D)
        shapes[p + 'dense/kernel'] = (1, D, R)
        shapes[p + 'skip/kernel'] = (1, D, S)
        if use_biases:
            shapes[p + 'dense/bias'] = (R,)
            shapes[p + 'skip/bias'] = (S,)
    out_dim = out_channels if scalar_input else quantization_channels
    shapes['wavenet/conv1d_1/kernel'] = (1, S, S)
    shapes['wavenet/conv1d_2/kernel'] = (1, S, out_dim)
    if use_biases:
        shapes['wavenet/conv1d_1/bias'] = (S,)
        shapes['wavenet/conv1d_2/bias'] = (out_dim,)
    return shapes


def make_weights(seed=1234, bias_scale=0.05, **model_kwargs):
    """Return {name: float32 ndarray}.  Deterministic in (seed, model_kwargs)."""
    rng = np.random.RandomState(seed)
    shapes = weight_shapes(**model_kwargs)
    state = {}
    for name, shp in shapes.items():
        if name.endswith('gc_embedding'):
            fan_in, fan_out = shp
            w = rng.randn(*shp) * np.sqrt(2.0 / (fan_in + fan_out))
        elif name.endswith('/bias'):
            w = rng.uniform(-bias_scale, bias_scale, shp)
        elif 'upsample' in name:
            # kernels of a trained upsampler are roughly "hold" filters; keep the signal O(1)
            w = rng.uniform(0.3, 0.7, shp)
        else:
            rf = int(np.prod(shp[:-2])) if len(shp) > 2 else 1
            fan_in, fan_out = rf * shp[-2], rf * shp[-1]
            lim = np.sqrt(6.0 / (fan_in + fan_out))
            w = rng.uniform(-lim, lim, shp)
        state[name] = w.astype(np.float32)
    if model_kwargs.get('scalar_input') and model_kwargs.get('use_biases'):
        nr = model_kwargs.get('out_channels', 30) // 3
        state['wavenet/conv1d_2/bias'][2 * nr:3 * nr] -= np.float32(3.0)
    return state


# Named configurations of BASELINE.md / SURVEY.md section 8(d).
def cfg1(batch_size=1):
    """BASELINE configs[0]: 10-layer mu-law WaveNet, unconditioned."""
    return dict(batch_size=batch_size, dilations=[1, 2, 4, 8, 16, 32, 64, 128, 256, 512], filter_width=2,
                residual_channels=32, dilation_channels=32, skip_channels=512, quantization_channels=256,
                use_biases=True, scalar_input=False, initial_filter_width=32, global_condition_channels=None,
                global_condition_cardinality=None, local_condition_channels=None, upsample_factor=None)


def cfg2(batch_size=8):
    """BASELINE configs[1]: 30-layer (3x10) R=D=128 MoL-10 mel-conditioned WaveNet."""
    return dict(batch_size=batch_size, dilations=[1, 2, 4, 8, 16, 32, 64, 128, 256, 512] * 3, filter_width=2,
                residual_channels=128, dilation_channels=128, skip_channels=512, quantization_channels=256,
                out_channels=30, use_biases=True, scalar_input=True, initial_filter_width=32,
                global_condition_channels=32, global_condition_cardinality=2, local_condition_channels=80,
                upsample_factor=[5, 5, 12])


def cfg_hparams_default(batch_size=1):
    """The reference's hparams.py:59-79 defaults: 50 layers, R=D=32, S=512, MoL, lc+gc."""
    return dict(batch_size=batch_size, dilations=[1, 2, 4, 8, 16, 32, 64, 128, 256, 512] * 5, filter_width=2,
                residual_channels=32, dilation_channels=32, skip_channels=512, quantization_channels=256,
                out_channels=30, use_biases=True, scalar_input=True, initial_filter_width=32,
                global_condition_channels=32, global_condition_cardinality=2, local_condition_channels=80,
                upsample_factor=[5, 5, 12])


def tiny_mol(batch_size=2):
    """Small mel-conditioned MoL model for fast parity tests."""
    return dict(batch_size=batch_size, dilations=[1, 2, 4, 1, 2, 4], filter_width=2, residual_channels=16,
                dilation_channels=16, skip_channels=32, quantization_channels=256, out_channels=30,
                use_biases=True, scalar_input=True, initial_filter_width=8, global_condition_channels=8,
                global_condition_cardinality=3, local_condition_channels=20, upsample_factor=[2, 3])


def tiny_mulaw(batch_size=2):
    """Small unconditioned mu-law model for fast parity tests."""
    return dict(batch_size=batch_size, dilations=[1, 2, 4, 8], filter_width=2, residual_channels=16,
                dilation_channels=16, skip_channels=64, quantization_channels=256, use_biases=True,
                scalar_input=False, initial_filter_width=32, global_condition_channels=None,
                global_condition_cardinality=None, local_condition_channels=None, upsample_factor=None)
