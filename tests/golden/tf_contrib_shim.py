# coding: utf-8
"""TEST INFRASTRUCTURE -- the tf.contrib.rnn / tf.contrib.seq2seq surface that the reference's tacotron/tacotron.py,
tacotron/rnn_wrappers.py and tacotron/helpers.py touch, on top of tf_numpy_shim (eager numpy).  With it the reference's
OWN Tacotron graph wiring runs here: embedding zero row, deepvoice speaker states, its AttentionWrapper (a modified copy of
TF's, with the manual-alignment override), DecoderPrenetWrapper, ConcatOutputAndAttentionWrapper, LocationSensitiveAttention,
TacoTestHelper, the cell stack and the post net.

The classes below ARE third-party TensorFlow code restated from its published definitions (TF 1.x python sources):
GRUCell, MultiRNNCell, OutputProjectionWrapper, ResidualWrapper, BahdanauAttention, BahdanauMonotonicAttention
(`_bahdanau_score`, `monotonic_attention(mode='parallel')`), BasicDecoder + dynamic_decode, TensorArray.  Variable scoping follows
TF's conventions (RNNCell.__call__ opens a scope named after the class in snake case; dynamic_decode opens "decoder"); the golden
script maps the resulting names onto the short names this repository uses for the decoder variables.
"""
import collections
import re
import sys
import types

import numpy as np

import tf_numpy_shim as tf

_t = tf._t


# ---- nest ----------------------------------------------------------------------------------------------------------------
def _is_seq(x):
    return isinstance(x, (list, tuple)) and not isinstance(x, str)


def flatten(s):
    if not _is_seq(s):
        return [s]
    out = []
    for e in s:
        out.extend(flatten(e))
    return out


def map_structure(fn, *structs):
    s0 = structs[0]
    if not _is_seq(s0):
        return fn(*structs)
    mapped = [map_structure(fn, *es) for es in zip(*structs)]
    if hasattr(s0, '_fields'):
        return type(s0)(*mapped)
    return type(s0)(mapped)


nest = types.SimpleNamespace(flatten=flatten, map_structure=map_structure, is_sequence=_is_seq)


# ---- small tf additions -----------------------------------------------------------------------------------------------------
class Placeholder(object):
    """tf.placeholder: the value is supplied with feed() before the graph-building code runs (eager emulation)."""

    def __init__(self, dtype, shape=None, name=None):
        self.dtype, self.name = dtype, name
        self.value = False if dtype is np.bool_ else None      # is_manual_attention defaults to False (synthesizer.py:138)

    def __getitem__(self, idx):
        return _t(np.asarray(self.value)[idx])


PLACEHOLDERS = {}
FEED = {}        # name -> value, consulted when the reference's code creates the placeholder (its feed_dict, synthesizer.py:129-160)


def placeholder(dtype, shape=None, name=None):
    p = Placeholder(dtype, shape, name)
    if name in FEED:
        p.value = FEED[name]
    PLACEHOLDERS[name] = p
    return p


def _val(x):
    return x.value if isinstance(x, Placeholder) else x


def cond(pred, true_fn, false_fn):
    return true_fn() if bool(np.asarray(_val(pred))) else false_fn()


def identity(x, name=None):
    return x


def assert_equal(a, b, message=None):
    assert int(a) == int(b), message
    return None


def matmul(a, b):
    return _t(np.matmul(np.asarray(a), np.asarray(b)))


def transpose(x, perm):
    return _t(np.transpose(np.asarray(x), perm))


def equal(a, b):
    return _t(np.asarray(a) == np.asarray(b))


def reduce_all(x, axis=None):
    return _t(np.all(np.asarray(x), axis=axis))


def zeros_initializer(*a, **k):
    return None


def TensorShape(dims):
    return list(dims)


class TensorArray(object):
    def __init__(self, dtype=None, size=0, dynamic_size=True, element_shape=None, **kw):
        self.items = {}

    def write(self, index, value):
        self.items[int(index)] = np.array(value)
        return self

    def stack(self):
        return _t(np.stack([self.items[i] for i in sorted(self.items)], 0))


def softsign(x):
    x = np.asarray(x)
    return _t(x / (np.abs(x) + 1))


class Dense(object):
    """tf.layers.Dense object: variables are created on the first call under the scope active THEN."""

    def __init__(self, units, activation=None, use_bias=True, name=None, dtype=None, **kw):
        self.units, self.activation, self.use_bias, self.name = units, activation, use_bias, name

    def __call__(self, x):
        return tf.layers.dense(x, self.units, activation=self.activation, use_bias=self.use_bias, name=self.name)


class Conv1D(object):
    def __init__(self, filters, kernel_size, padding='valid', use_bias=True, name=None, **kw):
        self.filters, self.k, self.padding, self.use_bias, self.name = filters, kernel_size[0] if isinstance(kernel_size, (tuple, list)) else kernel_size, padding, use_bias, name

    def __call__(self, x):
        return tf.layers.conv1d(x, self.filters, self.k, padding=self.padding, use_bias=self.use_bias, name=self.name)


# ---- RNN cells (tf.contrib.rnn) -------------------------------------------------------------------------------------------------
def _snake(name):
    return re.sub(r'(?<!^)(?=[A-Z])', '_', name).lower()


def _zero_state_tensors(state_size, batch_size, dtype):
    def one(s):
        dims = flatten(s) if _is_seq(s) else [s]
        return tf.zeros([int(batch_size)] + [int(d) for d in dims], np.float32)
    if _is_seq(state_size) and not isinstance(state_size, list):
        return map_structure(one, state_size)
    return one(state_size)


class RNNCell(object):
    def __init__(self, name=None, **kw):
        self._name = name or _snake(type(self).__name__)
        self._base_name = self._name

    @property
    def name(self):
        return self._name

    def __call__(self, inputs, state, scope=None):
        with tf.variable_scope(self._name):
            return self.call(inputs, state)

    def zero_state(self, batch_size, dtype):
        return _zero_state_tensors(self.state_size, batch_size, dtype)


class GRUCell(RNNCell):
    """[r, u] = sigmoid([x, h] Wg + bg); c = tanh([x, r*h] Wc + bc); h' = u*h + (1-u)*c  (variables gates/{kernel,bias},
    candidate/{kernel,bias})."""

    def __init__(self, num_units, **kw):
        super(GRUCell, self).__init__(name='gru_cell')
        self.num_units = int(num_units)

    @property
    def state_size(self):
        return self.num_units

    @property
    def output_size(self):
        return self.num_units

    def call(self, inputs, state):
        x, h = np.asarray(inputs), np.asarray(state)
        U = self.num_units
        Wg = np.asarray(tf._make_variable(tf._scoped('gates/kernel'), (x.shape[1] + U, 2 * U), True))
        bg = np.asarray(tf._make_variable(tf._scoped('gates/bias'), (2 * U,), True))
        Wc = np.asarray(tf._make_variable(tf._scoped('candidate/kernel'), (x.shape[1] + U, U), True))
        bc = np.asarray(tf._make_variable(tf._scoped('candidate/bias'), (U,), True))
        g = 1 / (1 + np.exp(-(np.concatenate([x, h], 1) @ Wg + bg)))
        r, u = g[:, :U], g[:, U:]
        c = np.tanh(np.concatenate([x, r * h], 1) @ Wc + bc)
        hn = _t((u * h + (1 - u) * c).astype(np.float32))
        return hn, hn

    def step(self, x, h, scope):      # used by tf_numpy_shim's bidirectional_dynamic_rnn
        saved = tf.S.scope
        tf.S.scope = scope.split('/')
        try:
            return np.asarray(self(x, h)[0])
        finally:
            tf.S.scope = saved


class MultiRNNCell(RNNCell):
    def __init__(self, cells, state_is_tuple=True):
        super(MultiRNNCell, self).__init__(name='multi_rnn_cell')
        self._cells = cells

    @property
    def state_size(self):
        return tuple(c.state_size for c in self._cells)

    @property
    def output_size(self):
        return self._cells[-1].output_size

    def zero_state(self, batch_size, dtype):
        return tuple(c.zero_state(batch_size, dtype) for c in self._cells)

    def call(self, inputs, state):
        cur, new_states = inputs, []
        for i, cell in enumerate(self._cells):
            with tf.variable_scope('cell_%d' % i):
                cur, ns = cell(cur, state[i])
            new_states.append(ns)
        return cur, tuple(new_states)


class OutputProjectionWrapper(RNNCell):
    def __init__(self, cell, output_size, activation=None):
        super(OutputProjectionWrapper, self).__init__(name='output_projection_wrapper')
        self._cell, self._output_size = cell, int(output_size)

    @property
    def state_size(self):
        return self._cell.state_size

    @property
    def output_size(self):
        return self._output_size

    def zero_state(self, batch_size, dtype):
        return self._cell.zero_state(batch_size, dtype)

    def call(self, inputs, state):
        output, res_state = self._cell(inputs, state)
        x = np.asarray(output)
        W = np.asarray(tf._make_variable(tf._scoped('kernel'), (x.shape[1], self._output_size), True))
        b = np.asarray(tf._make_variable(tf._scoped('bias'), (self._output_size,), True))
        return _t((x @ W + b).astype(np.float32)), res_state


class ResidualWrapper(RNNCell):
    def __init__(self, cell):
        super(ResidualWrapper, self).__init__(name='residual_wrapper')
        self._cell = cell

    @property
    def state_size(self):
        return self._cell.state_size

    @property
    def output_size(self):
        return self._cell.output_size

    def zero_state(self, batch_size, dtype):
        return self._cell.zero_state(batch_size, dtype)

    def __call__(self, inputs, state, scope=None):       # ResidualWrapper adds no scope of its own
        out, ns = self._cell(inputs, state)
        return _t(np.asarray(inputs) + np.asarray(out)), ns


# ---- attention (tf.contrib.seq2seq.python.ops.attention_wrapper) ---------------------------------------------------------------------
class AttentionMechanism(object):
    pass


class AttentionWrapperState(collections.namedtuple("AttentionWrapperState",
                                                   ("cell_state", "attention", "time", "alignments", "alignment_history", "attention_state"))):
    def clone(self, **kwargs):
        return super(AttentionWrapperState, self)._replace(**kwargs)


def _prepare_memory(memory, memory_sequence_length, check_inner_dims_defined=True):
    m = np.asarray(memory)
    if memory_sequence_length is None:
        return _t(m)
    mask = np.arange(m.shape[1])[None, :] < np.asarray(memory_sequence_length)[:, None]
    return _t(m * mask[:, :, None].astype(m.dtype))


def _maybe_mask_score(score, memory_sequence_length, score_mask_value):
    if memory_sequence_length is None:
        return score
    s = np.asarray(score)
    mask = np.arange(s.shape[1])[None, :] < np.asarray(memory_sequence_length)[:, None]
    return _t(np.where(mask, s, np.float32(score_mask_value)))


def _bahdanau_score(processed_query, keys, normalize):
    k = np.asarray(keys)
    num_units = k.shape[2]
    pq = np.asarray(processed_query)[:, None, :]
    v = np.asarray(tf.get_variable('attention_v', [num_units]))
    if normalize:
        g = np.asarray(tf.get_variable('attention_g', []))
        b = np.asarray(tf.get_variable('attention_b', [num_units]))
        normed_v = g * v * (1.0 / np.sqrt(np.sum(np.square(v))))
        return _t(np.sum(normed_v * np.tanh(k + pq + b), axis=2).astype(np.float32))
    return _t(np.sum(v * np.tanh(k + pq), axis=2).astype(np.float32))


def monotonic_attention(p_choose_i, previous_attention, mode):
    """tf.contrib.seq2seq.monotonic_attention, mode='parallel' (Raffel et al. 2017, closed form)."""
    if mode != 'parallel':
        raise NotImplementedError(mode)
    p, prev = np.asarray(p_choose_i), np.asarray(previous_attention)
    # safe_cumprod(1 - p, exclusive=True) = exp(cumsum(log(clip(1 - p, 1e-10, 1)), exclusive))
    lg = np.log(np.clip(1 - p, 1e-10, 1))
    cs = np.cumsum(lg, axis=1) - lg
    cumprod_1mp = np.exp(cs)
    return _t((p * cumprod_1mp * np.cumsum(prev / np.clip(cumprod_1mp, 1e-10, 1.), axis=1)).astype(np.float32))


def _monotonic_probability_fn(score, previous_alignments, sigmoid_noise, mode, seed=None):
    if sigmoid_noise:
        raise NotImplementedError('sigmoid_noise > 0')
    s = np.asarray(score)
    return monotonic_attention(1 / (1 + np.exp(-s)), previous_alignments, mode)


class _BaseAttentionMechanism(AttentionMechanism):
    def __init__(self, query_layer, memory, probability_fn, memory_sequence_length=None, memory_layer=None, check_inner_dims_defined=True,
                 score_mask_value=None, name=None):
        self._query_layer, self._memory_layer = query_layer, memory_layer
        if score_mask_value is None:
            score_mask_value = -np.inf
        self._probability_fn = lambda score, prev: probability_fn(_maybe_mask_score(score, memory_sequence_length, score_mask_value), prev)
        self._values = _prepare_memory(memory, memory_sequence_length)
        self._keys = self._memory_layer(self._values) if self._memory_layer else self._values     # variables under the CURRENT variable scope
        self._batch_size = int(np.asarray(self._keys).shape[0])
        self._alignments_size = int(np.asarray(self._keys).shape[1])
        self.dtype = np.float32

    memory_layer = property(lambda s: s._memory_layer)
    query_layer = property(lambda s: s._query_layer)
    values = property(lambda s: s._values)
    keys = property(lambda s: s._keys)
    batch_size = property(lambda s: s._batch_size)
    alignments_size = property(lambda s: s._alignments_size)
    state_size = property(lambda s: s._alignments_size)

    def initial_alignments(self, batch_size, dtype):
        return _zero_state_tensors(self._alignments_size, batch_size, dtype)

    def initial_state(self, batch_size, dtype):
        return self.initial_alignments(batch_size, dtype)


def _softmax_prob(score, _prev):
    return tf.nn.softmax(score)


class BahdanauAttention(_BaseAttentionMechanism):
    def __init__(self, num_units, memory, memory_sequence_length=None, normalize=False, probability_fn=None, score_mask_value=None,
                 dtype=None, name="BahdanauAttention"):
        if probability_fn is None:
            wrapped = _softmax_prob
        else:
            wrapped = lambda score, _: probability_fn(score)      # noqa: E731
        super(BahdanauAttention, self).__init__(query_layer=Dense(num_units, name="query_layer", use_bias=False),
                                                memory_layer=Dense(num_units, name="memory_layer", use_bias=False), memory=memory,
                                                probability_fn=wrapped, memory_sequence_length=memory_sequence_length,
                                                score_mask_value=score_mask_value, name=name)
        self._num_units, self._normalize, self._name = num_units, normalize, name

    def __call__(self, query, state):
        with tf.variable_scope("bahdanau_attention"):
            processed_query = self.query_layer(query) if self.query_layer else query
            score = _bahdanau_score(processed_query, self._keys, self._normalize)
        alignments = self._probability_fn(score, state)
        return alignments, alignments


class _BaseMonotonicAttentionMechanism(_BaseAttentionMechanism):
    def initial_alignments(self, batch_size, dtype):
        return tf.one_hot(np.zeros((int(batch_size),), np.int32), self._alignments_size, dtype=np.float32)


class BahdanauMonotonicAttention(_BaseMonotonicAttentionMechanism):
    def __init__(self, num_units, memory, memory_sequence_length=None, normalize=False, score_mask_value=None, sigmoid_noise=0.,
                 sigmoid_noise_seed=None, score_bias_init=0., mode="parallel", dtype=None, name="BahdanauMonotonicAttention"):
        fn = lambda score, prev: _monotonic_probability_fn(score, prev, sigmoid_noise, mode, sigmoid_noise_seed)      # noqa: E731
        super(BahdanauMonotonicAttention, self).__init__(query_layer=Dense(num_units, name="query_layer", use_bias=False),
                                                         memory_layer=Dense(num_units, name="memory_layer", use_bias=False), memory=memory,
                                                         probability_fn=fn, memory_sequence_length=memory_sequence_length,
                                                         score_mask_value=score_mask_value, name=name)
        self._num_units, self._normalize, self._name = num_units, normalize, name

    def __call__(self, query, state):
        with tf.variable_scope("bahdanau_monotonic_attention"):
            processed_query = self.query_layer(query) if self.query_layer else query
            score = _bahdanau_score(processed_query, self._keys, self._normalize)
            score_bias = tf.get_variable("attention_score_bias", [])
            score = _t(np.asarray(score) + np.asarray(score_bias))
        alignments = self._probability_fn(score, state)
        return alignments, alignments


class LuongAttention(AttentionMechanism):
    def __init__(self, *a, **k):
        raise NotImplementedError


# ---- decoder (tf.contrib.seq2seq) -------------------------------------------------------------------------------------------------------
class Helper(object):
    pass


class BasicDecoder(object):
    def __init__(self, cell, helper, initial_state, output_layer=None):
        self.cell, self.helper, self.initial_state = cell, helper, initial_state


def dynamic_decode(decoder, maximum_iterations=None, **kw):
    """while not all(finished) and t < maximum_iterations: outputs, state = cell(inputs, state); helper.next_inputs(...)"""
    outs = []
    with tf.variable_scope("decoder"):
        finished, inputs = decoder.helper.initialize()
        state = decoder.initial_state
        t = 0
        snap = dict(tf.S.counters)      # TF traces the loop body ONCE (tf.while_loop): default-name scopes are not re-uniquified per step
        while t < int(maximum_iterations) and not bool(np.all(np.asarray(finished))):
            tf.S.counters = dict(snap)
            cell_outputs, state = decoder.cell(inputs, state)
            sample_ids = decoder.helper.sample(time=t, outputs=cell_outputs, state=state)
            finished, inputs, state = decoder.helper.next_inputs(time=t, outputs=cell_outputs, state=state, sample_ids=sample_ids)
            outs.append(np.asarray(cell_outputs))
            t += 1
    return (_t(np.stack(outs, 1)), None), state, None


def install():
    """tensorflow + the contrib sub-modules the reference's tacotron package imports."""
    tf.install()
    for k, v in dict(placeholder=placeholder, cond=cond, identity=identity, assert_equal=assert_equal, matmul=matmul, transpose=transpose,
                     equal=equal, reduce_all=reduce_all, zeros_initializer=zeros_initializer, TensorShape=TensorShape, TensorArray=TensorArray,
                     bool=np.bool_).items():
        setattr(tf, k, v)
    tf.nn.softsign = staticmethod(softsign) if False else softsign
    tf.layers.Dense = Dense
    tf.layers.Conv1D = Conv1D
    tf.train.exponential_decay = lambda *a, **k: None
    rnn = types.ModuleType('tensorflow.contrib.rnn')
    rnn.__dict__.update(RNNCell=RNNCell, GRUCell=GRUCell, MultiRNNCell=MultiRNNCell, OutputProjectionWrapper=OutputProjectionWrapper,
                        ResidualWrapper=ResidualWrapper)
    aw = types.ModuleType('tensorflow.contrib.seq2seq.python.ops.attention_wrapper')
    aw.__dict__.update(_bahdanau_score=_bahdanau_score, _BaseAttentionMechanism=_BaseAttentionMechanism, BahdanauAttention=BahdanauAttention,
                       AttentionWrapper=None, AttentionWrapperState=AttentionWrapperState, AttentionMechanism=AttentionMechanism,
                       _BaseMonotonicAttentionMechanism=_BaseMonotonicAttentionMechanism, _maybe_mask_score=_maybe_mask_score,
                       _prepare_memory=_prepare_memory, _monotonic_probability_fn=_monotonic_probability_fn)
    s2s = types.ModuleType('tensorflow.contrib.seq2seq')
    s2s.__dict__.update(BasicDecoder=BasicDecoder, BahdanauAttention=BahdanauAttention, BahdanauMonotonicAttention=BahdanauMonotonicAttention,
                        LuongAttention=LuongAttention, Helper=Helper, dynamic_decode=dynamic_decode, monotonic_attention=monotonic_attention,
                        tile_batch=lambda x, multiplier: x, GreedyEmbeddingHelper=None)
    rci = types.ModuleType('tensorflow.python.ops.rnn_cell_impl')
    rci.__dict__.update(_zero_state_tensors=_zero_state_tensors, assert_like_rnncell=lambda name, cell: None)
    core = types.ModuleType('tensorflow.python.layers.core')
    core.Dense = Dense
    fw = types.ModuleType('tensorflow.contrib.framework')
    fw.nest = nest
    mods = {'tensorflow.contrib': types.ModuleType('tensorflow.contrib'), 'tensorflow.contrib.rnn': rnn, 'tensorflow.contrib.seq2seq': s2s,
            'tensorflow.contrib.seq2seq.python': types.ModuleType('x'), 'tensorflow.contrib.seq2seq.python.ops': types.ModuleType('x'),
            'tensorflow.contrib.seq2seq.python.ops.attention_wrapper': aw, 'tensorflow.python': types.ModuleType('x'),
            'tensorflow.python.ops': types.ModuleType('x'), 'tensorflow.python.ops.rnn_cell_impl': rci,
            'tensorflow.python.layers': types.ModuleType('x'), 'tensorflow.python.layers.core': core, 'tensorflow.contrib.framework': fw}
    mods['tensorflow.python.ops'].rnn_cell_impl = rci
    mods['tensorflow.python.layers'].core = core
    sys.modules.update(mods)
    tf.contrib.rnn = rnn
    tf.contrib.seq2seq = s2s
    tf.contrib.framework = fw
    # the eager bidirectional RNN of tf_numpy_shim drives cells through .step()
    tf.GRUCell = GRUCell
    return tf
