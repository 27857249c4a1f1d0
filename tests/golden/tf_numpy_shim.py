# coding: utf-8
"""TEST INFRASTRUCTURE -- a minimal eager numpy stand-in for the TensorFlow 1.x API surface that the reference's
wavenet/model.py, wavenet/mixture.py and wavenet/ops.py touch, so that THE REFERENCE'S OWN PYTHON runs in this container
(TensorFlow 1.x has no Python 3.12 wheel) and produces golden vectors for the oracle (tests/golden/make_reference_goldens.py).

What this pins: everything the reference's code decides -- graph wiring, variable names and shapes (tf.layers auto-numbering,
variable scopes), queue update order, slicing / alignment of the conditioning, the sampling formulas.  What it does NOT
pin: the arithmetic inside TensorFlow's kernels, which is restated here from the published op definitions (conv1d as a sum of
matmuls over taps, conv2d_transpose through torch.nn.functional.conv_transpose2d, softmax, ...), in numpy float32.

Semantics: ops execute eagerly on numpy arrays.  TF1 graph construction + repeated `sess.run` is emulated by calling the
reference's graph-building method once per step inside `graph_pass()`: variables (tf.Variable / tf.get_variable / layer
kernels) are created on the first pass and REUSED BY NAME afterwards; name uniquification (`dilation_queue`,
`dilation_queue_1`, ...; `conv1d`, `conv1d_1`, ...) restarts at every pass, so the i-th creation maps to the same variable.
`tf.scatter_update` mutates the variable, which is how the fast-generation queues keep their state between steps.
Initial values of trainable variables come from `set_initial_values({name: array})`; a name the reference asks for that is
missing from that dict raises, which is what checks SURVEY.md Appendix B.
"""
import contextlib
import sys
import types

import numpy as np

float32, float64, int32, int64 = np.float32, np.float64, np.int32, np.int64


class _Shape(list):
    def as_list(self):
        return list(self)


class Dim(int):
    """tf.Dimension: an int with `.value`."""

    @property
    def value(self):
        return int(self)


class Tensor(np.ndarray):
    """numpy array with the TF tensor surface the reference calls (`get_shape().as_list()`, `shape[i].value`)."""

    @property
    def shape(self):
        return tuple(Dim(d) for d in np.ndarray.shape.__get__(self))

    @shape.setter
    def shape(self, v):
        np.ndarray.shape.__set__(self, v)

    def get_shape(self):
        return _Shape(self.shape)


def _t(x, dtype=None):
    a = np.asarray(x, dtype=dtype)
    return a.view(Tensor)


class _State(object):
    def __init__(self):
        self.variables = {}          # full name -> Tensor (mutable storage)
        self.trainable = []          # names in creation order
        self.initial = {}
        self.scope = []              # variable-scope stack
        self.counters = {}           # per-pass unique-name counters
        self.uniforms = []           # queue for tf.random_uniform
        self.created_order = []
        self.resolver = None         # optional: TF-scoped name -> key of the supplied weight dict
        self.resolved = {}


S = _State()


def reset():
    S.__init__()


def set_initial_values(values):
    S.initial = {k: np.asarray(v, np.float32) for k, v in values.items()}


def push_uniforms(*arrays):
    S.uniforms.extend(np.asarray(a, np.float32) for a in arrays)


@contextlib.contextmanager
def graph_pass():
    """One emulated `sess.run`: unique-name counters restart so that creations map onto the existing variables."""
    S.counters = {}
    S.scope = []
    yield


def _scoped(name):
    return '/'.join(S.scope + [name])


def _unique(full):
    n = S.counters.get(full, 0)
    S.counters[full] = n + 1
    return full if n == 0 else '%s_%d' % (full, n)


def _make_variable(full, shape, trainable, init=None):
    if full in S.variables:
        v = S.variables[full]
        if tuple(v.shape) != tuple(shape):
            raise ValueError('variable %s re-created with shape %s (was %s)' % (full, tuple(shape), v.shape))
        return v
    if init is None:
        key = S.resolver(full) if S.resolver else full
        if key not in S.initial:
            raise KeyError('the reference created trainable variable %r %s (-> %r), which the supplied weights do not contain' % (full, tuple(shape), key))
        S.resolved[full] = key
        init = S.initial[key]
        if tuple(init.shape) != tuple(shape):
            raise ValueError('variable %s: reference shape %s, supplied %s' % (full, tuple(shape), init.shape))
    v = _t(np.array(init, np.float32, copy=True))
    v.name = full + ':0'
    S.variables[full] = v
    S.created_order.append(full)
    if trainable:
        S.trainable.append(full)
    return v


# ---- scopes -----------------------------------------------------------------------------------------------------------
AUTO_REUSE = 'auto_reuse'


@contextlib.contextmanager
def variable_scope(name_or_scope, default_name=None, values=None, reuse=None):
    """tf.variable_scope(name_or_scope, default_name=None, values=None, ..., reuse=None): with name None the default name is
    uniquified inside the enclosing scope."""
    if not isinstance(default_name, (str, type(None))):      # reference call sites pass reuse= by keyword only
        default_name = None
    name = name_or_scope if name_or_scope is not None else _unique(_scoped(default_name)).rsplit('/', 1)[-1]
    S.scope.append(name)
    try:
        yield name
    finally:
        S.scope.pop()


@contextlib.contextmanager
def name_scope(name, *a, **k):
    yield


@contextlib.contextmanager
def control_dependencies(_):
    yield


# ---- variables --------------------------------------------------------------------------------------------------------
def Variable(initial_value=None, name=None, trainable=True, dtype=None):
    full = _unique(_scoped(name or 'Variable'))
    return _make_variable(full, np.shape(initial_value), trainable, init=None if trainable else initial_value)


def get_variable(name, shape=None, dtype=None, initializer=None, trainable=True):
    return _make_variable(_scoped(name), tuple(int(d) for d in shape), trainable)


def trainable_variables():
    return [S.variables[n] for n in S.trainable]


def variables_initializer(var_list):
    def run():
        for v in var_list:
            v[...] = 0
    return run


def scatter_update(ref, indices, updates):
    ref[np.asarray(indices)] = np.asarray(updates, ref.dtype)
    return ref


# ---- array ops --------------------------------------------------------------------------------------------------------
def zeros(shape, dtype=np.float32):
    return _t(np.zeros([int(d) for d in shape], dtype))


def range(*a):          # noqa: A001 (mirrors tf.range)
    return _t(np.arange(*a))


def shape(x):
    return np.array(np.shape(x), np.int64)


def reshape(x, shp):
    return _t(np.reshape(np.asarray(x), [int(s) for s in shp]))


def concat(values, axis):
    return _t(np.concatenate([np.asarray(v) for v in values], axis=axis))


def slice(x, begin, size):   # noqa: A001
    x = np.asarray(x)
    idx = tuple(np.s_[int(b):(None if int(s) == -1 else int(b) + int(s))] for b, s in zip(begin, size))
    return _t(x[idx])


def cast(x, dtype):
    return _t(np.asarray(x).astype(dtype))


def to_float(x):
    return cast(x, np.float32)


def to_int32(x):
    # tf.to_int32 truncates toward zero
    return _t(np.trunc(np.asarray(x)).astype(np.int32))


def expand_dims(x, axis):
    ax = axis[0] if isinstance(axis, (list, tuple)) else axis
    return _t(np.expand_dims(np.asarray(x), ax))


def squeeze(x, axis=None):
    ax = tuple(axis) if isinstance(axis, (list, tuple)) else axis
    return _t(np.squeeze(np.asarray(x), axis=ax))


def tile(x, multiples):
    a = np.asarray(x)
    if a.dtype == np.float64 and not isinstance(x, np.ndarray):
        a = a.astype(np.float32)           # a Python float constant is a float32 tensor in TF
    return _t(np.tile(a, [int(m) for m in multiples]))


def one_hot(indices, depth, dtype=np.float32):
    idx = np.asarray(indices).astype(np.int64)
    out = np.zeros(idx.shape + (int(depth),), dtype)
    np.put_along_axis(out, idx[..., None], 1, axis=-1)
    return _t(out)


def where(cond, a, b):
    return _t(np.where(np.asarray(cond), np.asarray(a), np.asarray(b)))


def argmax(x, axis):
    return _t(np.argmax(np.asarray(x), axis=axis))


# ---- math -------------------------------------------------------------------------------------------------------------
def _un(fn):
    return lambda x, *a, **k: _t(fn(np.asarray(x)))


tanh = _un(np.tanh)
exp = _un(np.exp)
log = _un(np.log)
log1p = _un(np.log1p)
abs = _un(np.abs)          # noqa: A001
sign = _un(np.sign)


def sigmoid(x):
    x = np.asarray(x)
    return _t((1 / (1 + np.exp(-x))).astype(x.dtype))


def maximum(a, b):
    return _t(np.maximum(np.asarray(a), np.asarray(b, dtype=np.asarray(a).dtype)))


def minimum(a, b):
    return _t(np.minimum(np.asarray(a), np.asarray(b, dtype=np.asarray(a).dtype)))


def _ax(axis):
    return tuple(axis) if isinstance(axis, (list, tuple)) else axis


def reduce_sum(x, axis=None, keepdims=False):
    return _t(np.sum(np.asarray(x), axis=_ax(axis), keepdims=keepdims))


def reduce_max(x, axis=None, keepdims=False):
    return _t(np.max(np.asarray(x), axis=axis, keepdims=keepdims))


def reduce_mean(x, axis=None):
    return _t(np.mean(np.asarray(x), axis=axis, dtype=np.asarray(x).dtype))


def add_n(xs):
    out = np.asarray(xs[0])
    for x in xs[1:]:
        out = out + np.asarray(x)
    return _t(out)


def random_uniform(shp, minval=0, maxval=1, dtype=np.float32):
    """The caller queues the draws (push_uniforms) so that reference and oracle consume the same random numbers."""
    if not S.uniforms:
        raise RuntimeError('tf.random_uniform called but no uniforms were queued')
    u = S.uniforms.pop(0)
    if tuple(u.shape) != tuple(int(s) for s in shp):
        raise ValueError('queued uniforms have shape %s, the reference asked for %s' % (u.shape, tuple(shp)))
    assert u.min() >= minval and u.max() <= maxval
    return _t(u)


class _NN(object):
    relu = staticmethod(lambda x: _t(np.maximum(np.asarray(x), 0)))
    sigmoid = staticmethod(sigmoid)

    @staticmethod
    def softplus(x):
        x = np.asarray(x)
        return _t(np.logaddexp(0, x).astype(x.dtype))

    @staticmethod
    def softmax(x, axis=-1):
        x = np.asarray(x)
        e = np.exp(x - x.max(axis=axis, keepdims=True))
        return _t(e / e.sum(axis=axis, keepdims=True))

    @staticmethod
    def log_softmax(x, axis=-1):
        x = np.asarray(x)
        m = x - x.max(axis=axis, keepdims=True)
        return _t(m - np.log(np.exp(m).sum(axis=axis, keepdims=True)))

    @staticmethod
    def embedding_lookup(table, ids):
        return _t(np.asarray(table)[np.asarray(ids).astype(np.int64)])

    @staticmethod
    def softmax_cross_entropy_with_logits_v2(logits=None, labels=None):
        ls = _NN.log_softmax(logits)
        return _t(-(np.asarray(labels) * ls).sum(-1))

    @staticmethod
    def l2_loss(v):
        return _t(np.sum(np.asarray(v) ** 2) / 2)


nn = _NN()


def split(x, num, axis):
    return [_t(a) for a in np.split(np.asarray(x), num, axis=axis)]


def truncated_normal_initializer(**k):
    return None


def constant_initializer(*a, **k):
    return None


class GRUCell(object):
    """tf.contrib.rnn.GRUCell: [r, u] = sigmoid([x, h] Wg + bg); c = tanh([x, r*h] Wc + bc); h' = u*h + (1-u)*c.
    Variables `<scope>/gru_cell/{gates,candidate}/{kernel,bias}` (third-party arithmetic, restated)."""

    def __init__(self, num_units):
        self.num_units = int(num_units)

    def step(self, x, h, scope):
        U = self.num_units
        Wg = np.asarray(_make_variable(scope + '/gru_cell/gates/kernel', (x.shape[1] + U, 2 * U), True))
        bg = np.asarray(_make_variable(scope + '/gru_cell/gates/bias', (2 * U,), True))
        Wc = np.asarray(_make_variable(scope + '/gru_cell/candidate/kernel', (x.shape[1] + U, U), True))
        bc = np.asarray(_make_variable(scope + '/gru_cell/candidate/bias', (U,), True))
        g = 1 / (1 + np.exp(-(np.concatenate([x, h], 1) @ Wg + bg)))
        r, u = g[:, :U], g[:, U:]
        c = np.tanh(np.concatenate([x, r * h], 1) @ Wc + bc)
        return (u * h + (1 - u) * c).astype(np.float32)


def _bidirectional_dynamic_rnn(cell_fw, cell_bw, inputs, sequence_length=None, initial_state_fw=None, initial_state_bw=None, dtype=None):
    """tf.nn.bidirectional_dynamic_rnn: outputs are zero past sequence_length and the state is carried through; the backward
    direction runs over the reversed valid part (tf.reverse_sequence)."""
    x = np.asarray(inputs)
    N, T, _ = x.shape
    L = np.full((N,), T) if sequence_length is None else np.asarray(sequence_length)
    scope = _scoped('bidirectional_rnn')
    outs = []
    for cell, d, init in ((cell_fw, 'fw', initial_state_fw), (cell_bw, 'bw', initial_state_bw)):
        h = np.zeros((N, cell.num_units), np.float32) if init is None else np.asarray(init, np.float32).copy()
        out = np.zeros((N, T, cell.num_units), np.float32)
        order = np.arange(T) if d == 'fw' else np.arange(T)[::-1]
        for t in order:
            live = (t < L)[:, None]
            hn = cell.step(x[:, t], h, scope + '/' + d)
            h = np.where(live, hn, h)
            out[:, t] = np.where(live, hn, 0)
        outs.append(_t(out))
    return tuple(outs), None


_NN.bidirectional_dynamic_rnn = staticmethod(_bidirectional_dynamic_rnn)
_NN.tanh = staticmethod(lambda x: tanh(x))


# ---- tf.layers --------------------------------------------------------------------------------------------------------
def _layer_scope(name, default):
    if name is not None:
        return _scoped(name)
    return _unique(_scoped(default))


class _Layers(object):
    @staticmethod
    def conv1d(inputs, filters, kernel_size, padding='valid', dilation_rate=1, use_bias=True, name=None, strides=1, activation=None):
        x = np.asarray(inputs)
        scope = _layer_scope(name, 'conv1d')
        k = int(kernel_size)
        W = _make_variable(scope + '/kernel', (k, x.shape[2], int(filters)), True)
        d = int(dilation_rate)
        if padding.lower() == 'same' and k != 1:
            # TF 'SAME', stride 1: total padding k-1, the smaller half in front
            assert d == 1
            x = np.pad(x, ((0, 0), ((k - 1) // 2, (k - 1) - (k - 1) // 2), (0, 0)))
        t_out = x.shape[1] - d * (k - 1)
        out = np.zeros((x.shape[0], t_out, int(filters)), np.float32)
        for j in np.arange(k):
            out = out + x[:, j * d:j * d + t_out, :] @ np.asarray(W)[j]
        if use_bias:
            out = out + np.asarray(_make_variable(scope + '/bias', (int(filters),), True))
        out = _t(out.astype(np.float32))
        return activation(out) if activation is not None else out

    @staticmethod
    def dense(inputs, units, activation=None, use_bias=True, name=None, bias_initializer=None, kernel_initializer=None):
        x = np.asarray(inputs)
        scope = _layer_scope(name, 'dense')
        W = _make_variable(scope + '/kernel', (x.shape[-1], int(units)), True)
        out = x @ np.asarray(W)
        if use_bias:
            out = out + np.asarray(_make_variable(scope + '/bias', (int(units),), True))
        out = _t(out.astype(np.float32))
        return activation(out) if activation is not None else out

    @staticmethod
    def dropout(inputs, rate=0.5, training=False, name=None):
        if training and rate:
            raise NotImplementedError('dropout in training mode')
        return inputs

    @staticmethod
    def batch_normalization(inputs, training=False, epsilon=1e-3, name=None):
        """Inference mode: (x - moving_mean) / sqrt(moving_variance + eps) * gamma + beta, eps = 1e-3 (tf.layers default)."""
        if training:
            raise NotImplementedError('batch_normalization in training mode')
        x = np.asarray(inputs)
        scope = _layer_scope(name, 'batch_normalization')
        c = x.shape[-1]
        g, b, m, v = (np.asarray(_make_variable(scope + '/' + n, (c,), True)) for n in ('gamma', 'beta', 'moving_mean', 'moving_variance'))
        return _t(((x - m) / np.sqrt(v + np.float32(epsilon)) * g + b).astype(np.float32))

    @staticmethod
    def max_pooling1d(inputs, pool_size, strides, padding='valid'):
        x = np.asarray(inputs)
        if int(pool_size) != 2 or int(strides) != 1 or padding.lower() != 'same':
            raise NotImplementedError('only max_pooling1d(2, 1, same) (modules.py:38)')
        # 'SAME' pads one step at the END (with -inf): out[t] = max(x[t], x[t+1]), last step alone
        nxt = np.concatenate([x[:, 1:], np.full_like(x[:, :1], -np.inf)], axis=1)
        return _t(np.maximum(x, nxt))

    @staticmethod
    def conv2d_transpose(inputs, filters, kernel_size, strides, padding='same', use_bias=True, name=None):
        """NHWC, filters=1, kernel (F, fw), strides (F, 1), 'same' (model.py:107-108).  TF's 'same' transposed convolution with
        an even width-2 kernel crops the END of the full output, i.e. keeps the first W columns."""
        import torch
        x = np.asarray(inputs)
        if filters != 1 or x.shape[3] != 1 or padding.lower() != 'same' or use_bias or strides[1] != 1 or strides[0] != kernel_size[0]:
            raise NotImplementedError('only the upsampling configuration of model.py:107-108')
        scope = _layer_scope(name, 'conv2d_transpose')
        K = _make_variable(scope + '/kernel', (kernel_size[0], kernel_size[1], 1, 1), True)
        xt = torch.from_numpy(np.ascontiguousarray(x[..., 0]))[:, None]                       # N,1,H,W
        kt = torch.from_numpy(np.ascontiguousarray(np.asarray(K)[:, :, 0, 0]))[None, None]     # 1,1,F,fw
        y = torch.nn.functional.conv_transpose2d(xt, kt, stride=(int(strides[0]), 1))[..., :x.shape[2]]
        return _t(y[:, 0].numpy()[..., None].astype(np.float32))


layers = _Layers()


# ---- things the reference touches but that do nothing here ---------------------------------------------------------------
class _Summary(object):
    @staticmethod
    def scalar(*a, **k):
        return None


summary = _Summary()


class _Train(object):
    class ExponentialMovingAverage(object):
        def __init__(self, decay):
            self.decay = decay

    AdamOptimizer = MomentumOptimizer = RMSPropOptimizer = object


train = _Train()

contrib = types.SimpleNamespace(layers=types.SimpleNamespace(xavier_initializer=lambda **k: None), rnn=types.SimpleNamespace(GRUCell=GRUCell))


def install():
    """Registers this module as `tensorflow` (only in the process that generates the golden vectors)."""
    me = sys.modules[__name__]
    sys.modules['tensorflow'] = me
    # sub-modules the reference imports by path (tacotron/modules.py:5-7); only GRUCell is used by the functions run here
    for name, attrs in (('tensorflow.contrib', {}), ('tensorflow.contrib.rnn', {'GRUCell': GRUCell}), ('tensorflow.python', {}),
                        ('tensorflow.python.layers', {'core': None}), ('tensorflow.contrib.seq2seq', {}), ('tensorflow.contrib.seq2seq.python', {}),
                        ('tensorflow.contrib.seq2seq.python.ops', {}),
                        ('tensorflow.contrib.seq2seq.python.ops.attention_wrapper',
                         dict.fromkeys(['_bahdanau_score', '_BaseAttentionMechanism', 'BahdanauAttention', 'AttentionWrapper', 'AttentionWrapperState']))):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules.setdefault(name, m)
    return me
