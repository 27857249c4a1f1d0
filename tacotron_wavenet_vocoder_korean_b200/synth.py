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


def make_weights(seed=1234, bias_scale=0.05, init='benchmark', **model_kwargs):
    """Return {name: float32 ndarray}.  Deterministic in (seed, model_kwargs).

    init='benchmark': synthetic weights of a plausible *trained* net (upsample kernels ~ hold filters U(0.3, 0.7), small random biases,
    log-scale bias shifted) for the generation benchmarks and parity tests.  init='train': what tf.global_variables_initializer gives
    the reference graph (train_vocoder.py:129-130): Glorot-uniform for every kernel including the conv2d_transpose upsamplers, zero
    biases, Xavier-normal gc_embedding; pass a fresh seed per run."""
    if init not in ('benchmark', 'train'):
        raise ValueError("init must be 'benchmark' or 'train'")
    train = init == 'train'
    if train:
        bias_scale = 0.0
    rng = np.random.RandomState(seed)
    shapes = weight_shapes(**model_kwargs)
    state = {}
    for name, shp in shapes.items():
        if name.endswith('gc_embedding'):
            fan_in, fan_out = shp
            w = rng.randn(*shp) * np.sqrt(2.0 / (fan_in + fan_out))
        elif name.endswith('/bias'):
            w = rng.uniform(-bias_scale, bias_scale, shp)
        elif 'upsample' in name and not train:
            # kernels of a trained upsampler are roughly "hold" filters; keep the signal O(1)
            w = rng.uniform(0.3, 0.7, shp)
        else:
            rf = int(np.prod(shp[:-2])) if len(shp) > 2 else 1
            fan_in, fan_out = rf * shp[-2], rf * shp[-1]
            lim = np.sqrt(6.0 / (fan_in + fan_out))
            w = rng.uniform(-lim, lim, shp)
        state[name] = w.astype(np.float32)
    if model_kwargs.get('scalar_input') and model_kwargs.get('use_biases') and not train:
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


def tiny_train(batch_size=3):
    """Small MoL model for training-step parity (all channel counts multiples of 8, as the bf16 GEMMs want)."""
    return dict(batch_size=batch_size, dilations=[1, 2, 4, 1, 2, 4], filter_width=2, residual_channels=16,
                dilation_channels=32, skip_channels=64, quantization_channels=256, out_channels=30,
                use_biases=True, scalar_input=True, initial_filter_width=8, global_condition_channels=8,
                global_condition_cardinality=3, local_condition_channels=24, upsample_factor=[2, 3])


def tiny_mulaw(batch_size=2):
    """Small unconditioned mu-law model for fast parity tests."""
    return dict(batch_size=batch_size, dilations=[1, 2, 4, 8], filter_width=2, residual_channels=16,
                dilation_channels=16, skip_channels=64, quantization_channels=256, use_biases=True,
                scalar_input=False, initial_filter_width=32, global_condition_channels=None,
                global_condition_cardinality=None, local_condition_channels=None, upsample_factor=None)


# ------------------------------------------------------------------------------------------------------
# Tacotron.  Variable names follow the reference's TF scopes where they are
# fixed by its own code (tacotron.py:46, modules.py:15-96: model/inference/{embedding, speaker_embedding,
# dense*, prenet, encoder_cbhg, post_cbhg}); the decoder-cell names come from tf.contrib wrappers and are
# not pinned by the reference, so short names under model/inference/decoder/ are used.
TACO_HP = dict(
    num_symbols=80, embedding_size=256, speaker_embedding_size=16, model_type='deepvoice',
    enc_prenet_sizes=[256, 128], enc_bank_size=16, enc_bank_channel_size=128, enc_maxpool_width=2,
    enc_highway_depth=4, enc_rnn_size=128, enc_proj_sizes=[128, 128], enc_proj_width=3,
    attention_type='bah_mon_norm', attention_size=256, attention_state_size=256,
    dec_layer_num=2, dec_rnn_size=256, dec_prenet_sizes=[256, 128],
    post_bank_size=8, post_bank_channel_size=128, post_maxpool_width=2, post_highway_depth=4,
    post_rnn_size=128, post_proj_sizes=[256, 80], post_proj_width=3,
    reduction_factor=5, max_iters=200, num_mels=80, num_freq=1025,
)


def taco_tiny(**over):
    """Small Tacotron for fast parity tests (every code path of the full model, small widths)."""
    hp = dict(TACO_HP)
    hp.update(embedding_size=32, speaker_embedding_size=8, enc_prenet_sizes=[32, 16], enc_bank_size=4,
              enc_bank_channel_size=16, enc_rnn_size=16, enc_proj_sizes=[16, 16], attention_size=32,
              attention_state_size=32, dec_rnn_size=32, dec_prenet_sizes=[32, 16], post_bank_size=3,
              post_bank_channel_size=16, post_rnn_size=16, post_proj_sizes=[32, 20], reduction_factor=2,
              max_iters=12, num_mels=20, num_freq=65)
    hp.update(over)
    return hp


def _gru_shapes(shapes, scope, n_in, units):
    shapes[scope + '/gates/kernel'] = (n_in + units, 2 * units)
    shapes[scope + '/gates/bias'] = (2 * units,)
    shapes[scope + '/candidate/kernel'] = (n_in + units, units)
    shapes[scope + '/candidate/bias'] = (units,)


def _cbhg_shapes(shapes, scope, n_in, K, bank_ch, proj_sizes, proj_width, depth, rnn_size):
    def conv(s, k, ci, co):
        shapes[s + '/conv1d/kernel'] = (k, ci, co)
        shapes[s + '/conv1d/bias'] = (co,)
        for n in ('gamma', 'beta', 'moving_mean', 'moving_variance'):
            shapes[s + '/batch_normalization/' + n] = (co,)
    for k in range(1, K + 1):
        conv('%s/conv_bank/conv1d_%d' % (scope, k), k, n_in, bank_ch)
    ci = K * bank_ch
    for i, co in enumerate(proj_sizes):
        conv('%s/proj_%d' % (scope, i + 1), proj_width, ci, co)
        ci = co
    if ci != rnn_size:
        shapes[scope + '/dense/kernel'] = (ci, rnn_size)
        shapes[scope + '/dense/bias'] = (rnn_size,)
    for i in range(depth):
        for g in ('H', 'T'):
            shapes['%s/highway_%d/%s/kernel' % (scope, i + 1, g)] = (rnn_size, rnn_size)
            shapes['%s/highway_%d/%s/bias' % (scope, i + 1, g)] = (rnn_size,)
    for d in ('fw', 'bw'):
        _gru_shapes(shapes, '%s/bidirectional_rnn/%s/gru_cell' % (scope, d), rnn_size, rnn_size)


def taco_weight_shapes(hp, num_speakers):
    """Ordered {tf_variable_name: shape} of the inference graph (tacotron.py:36-219)."""
    P = 'model/inference/'
    s = {}
    E = hp['embedding_size']
    s['embedding'] = (hp['num_symbols'], E)
    nd = 0
    if num_speakers > 1:
        se = hp['speaker_embedding_size']
        s['speaker_embedding'] = (num_speakers, se)
        for width in [hp['enc_prenet_sizes'][-1], 2 * hp['enc_rnn_size'], hp['attention_state_size']] + \
                [hp['dec_rnn_size']] * hp['dec_layer_num']:
            n = 'dense' if nd == 0 else 'dense_%d' % nd
            s[n + '/kernel'] = (se, width)
            s[n + '/bias'] = (width,)
            nd += 1
    ci = E
    for i, co in enumerate(hp['enc_prenet_sizes']):
        s['prenet/dense_%d/kernel' % (i + 1)] = (ci, co)
        s['prenet/dense_%d/bias' % (i + 1)] = (co,)
        ci = co
    assert hp['enc_proj_sizes'][-1] == ci, "encoder CBHG residual needs proj_sizes[-1] == prenet width"
    _cbhg_shapes(s, 'encoder_cbhg', ci, hp['enc_bank_size'], hp['enc_bank_channel_size'], hp['enc_proj_sizes'],
                 hp['enc_proj_width'], hp['enc_highway_depth'], hp['enc_rnn_size'])
    mem = 2 * hp['enc_rnn_size']
    A = hp['attention_size']
    s['memory_layer/kernel'] = (mem, A)
    D = 'decoder/'
    nm = hp['num_mels']
    ci = nm
    for i, co in enumerate(hp['dec_prenet_sizes']):
        s[D + 'decoder_prenet/dense_%d/kernel' % (i + 1)] = (ci, co)
        s[D + 'decoder_prenet/dense_%d/bias' % (i + 1)] = (co,)
        ci = co
    H = hp['attention_state_size']
    _gru_shapes(s, D + 'attention_cell/gru_cell', ci + mem, H)
    s[D + 'attention/query_layer/kernel'] = (H, A)
    at = hp['attention_type']
    if at in ('bah_mon_norm', 'bah_mon'):
        s[D + 'attention/attention_v'] = (A,)
        if at == 'bah_mon_norm':
            s[D + 'attention/attention_g'] = ()
            s[D + 'attention/attention_b'] = (A,)
        s[D + 'attention/attention_score_bias'] = ()
    elif at == 'loc_sen':
        s[D + 'attention/attention_variable'] = (A,)
        s[D + 'attention/attention_bias'] = (A,)
        s[D + 'attention/location_features_convolution/kernel'] = (31, 1, 32)
        s[D + 'attention/location_features_convolution/bias'] = (32,)
        s[D + 'attention/location_features_layer/kernel'] = (32, A)
    else:
        raise ValueError("unsupported attention_type %r" % at)
    R = hp['dec_rnn_size']
    s[D + 'concat_projection/kernel'] = (H + mem, R)
    s[D + 'concat_projection/bias'] = (R,)
    for i in range(hp['dec_layer_num']):
        _gru_shapes(s, D + 'cell_%d/gru_cell' % (i + 1), R, R)
    s[D + 'output_projection/kernel'] = (R, nm * hp['reduction_factor'])
    s[D + 'output_projection/bias'] = (nm * hp['reduction_factor'],)
    assert hp['post_proj_sizes'][-1] == nm, "post CBHG residual needs proj_sizes[-1] == num_mels"
    _cbhg_shapes(s, 'post_cbhg', nm, hp['post_bank_size'], hp['post_bank_channel_size'], hp['post_proj_sizes'],
                 hp['post_proj_width'], hp['post_highway_depth'], hp['post_rnn_size'])
    n = 'dense' if nd == 0 else 'dense_%d' % nd
    s[n + '/kernel'] = (2 * hp['post_rnn_size'], hp['num_freq'])
    s[n + '/bias'] = (hp['num_freq'],)
    return {P + k: v for k, v in s.items()}


def make_taco_weights(hp, num_speakers, seed=4321):
    """Seeded synthetic Tacotron weights: Glorot-uniform kernels, TF-default-like special cases (GRU gate bias 1,
    highway T bias -1, batch-norm statistics near identity but not identity), small random biases so that every
    bias path is exercised, attention score bias -1.5 so that the monotonic alignment advances slowly."""
    rng = np.random.RandomState(seed)
    out = {}
    for name, shp in taco_weight_shapes(hp, num_speakers).items():
        if name.endswith('embedding'):
            w = np.clip(rng.randn(*shp), -2, 2) * 0.5
        elif name.endswith('/gamma'):
            w = rng.uniform(0.8, 1.2, shp)
        elif name.endswith('/moving_variance'):
            w = rng.uniform(0.5, 1.5, shp)
        elif name.endswith('/beta') or name.endswith('/moving_mean'):
            w = rng.uniform(-0.1, 0.1, shp)
        elif name.endswith('attention_g'):
            w = np.sqrt(1.0 / hp['attention_size']) * 4.0
        elif name.endswith('attention_score_bias'):
            w = -1.5
        elif name.endswith('/bias') or name.endswith('attention_b') or name.endswith('attention_bias'):
            w = rng.uniform(-0.05, 0.05, shp)
            if name.endswith('gates/bias'):
                w += 1.0
            if '/T/bias' in name:
                w -= 1.0
        elif len(shp) == 1:
            lim = np.sqrt(6.0 / (1 + shp[0]))
            w = rng.uniform(-lim, lim, shp)
        else:
            rf = int(np.prod(shp[:-2])) if len(shp) > 2 else 1
            lim = np.sqrt(6.0 / (rf * shp[-2] + rf * shp[-1]))
            w = rng.uniform(-lim, lim, shp)
        out[name] = np.asarray(w, dtype=np.float32)
    return out
