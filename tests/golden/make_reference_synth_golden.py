# coding: utf-8
"""Generates tests/golden/ref_synth_main.npz by running THE REFERENCE'S OWN synthesizer.py -- `Synthesizer.load` (placeholders,
load_hparams from params.json, create_model + initialize, Saver.restore) and `Synthesizer.synthesize` (text_to_sequence,
_prepare_inputs, input_lengths, the feed_dict, plot_graph_and_save_audio with the attention trimming, the mel .npy it writes for
generate.py --mel) -- unmodified, on the numpy TensorFlow stand-ins.  `tf.Session.run(fetches, feed_dict)` re-traces
`Tacotron.initialize` with the fed values (same emulation as make_reference_generate_golden.py).  Plotting and Griffin-Lim are
replaced by recorders; jamo / unidecode / inflect / matplotlib / librosa are import stubs.

    python tests/golden/make_reference_synth_golden.py        (build container only)
"""
import contextlib
import glob
import io
import json
import os
import sys
import tempfile
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get('REFERENCE_ROOT', '/root/reference')
sys.path.insert(0, HERE)
import make_reference_audio_golden as mra                  # noqa: E402
mra.install_stubs()
import make_reference_taco_full_golden as mtf              # noqa: E402  (tf + contrib stand-ins, the reference tacotron package, NAME_MAP)
import make_reference_text_golden as mtext                 # noqa: E402
from make_reference_generate_golden import Placeholder as _ArrayPlaceholder   # noqa: E402

tf, contrib = mtf.tf, mtf.contrib


class ArrayPlaceholder(_ArrayPlaceholder):
    def __init__(self, dtype, shape=None, name=None):
        _ArrayPlaceholder.__init__(self, dtype, shape, name)
        self.name = name


TEXTS = ['존경하는 국민 여러분', '오늘은 날씨가 좋습니다', '네']
SPEAKERS = [0, 1, 1]
STATE = {}
STATE_FEED = {}


class Session(object):
    def __init__(self, config=None):
        pass

    def run(self, fetches, feed_dict=None):
        if callable(fetches) or fetches is None:
            return None
        for p, v in (feed_dict or {}).items():
            if isinstance(p, contrib.Placeholder):
                p.value = v
                contrib.FEED[p.name] = v
            else:
                p.value = np.asarray(v).astype(p.dtype)
                STATE_FEED[p.name] = np.array(p.value)
        model, args, kw = STATE['init']
        with tf.graph_pass():
            with tf.variable_scope('model'):
                STATE['orig_initialize'](model, *args, **kw)
        return [np.array(getattr(model, STATE['names'][id(f)])) for f in fetches]

    def close(self):
        pass


def main():
    sys.path.insert(1, ROOT)
    from tacotron_wavenet_vocoder_korean_b200 import synth
    # third-party import stubs
    jm = types.ModuleType('jamo')
    jm.hangul_to_jamo = mtext.hangul_to_jamo
    jm.h2j = lambda s: ''.join(mtext.hangul_to_jamo(s))
    jm.j2hcj = lambda s: s
    jm.j2h = lambda lead, vowel, tail=None: chr(0xAC00 + (ord(lead) - 0x1100) * 588 + (ord(vowel) - 0x1161) * 28 + ((ord(tail) - 0x11A7) if tail else 0))
    jm.__path__ = []
    jmj = types.ModuleType('jamo.jamo')
    jmj._jamo_char_to_hcj = lambda c: c
    un = types.ModuleType('unidecode')
    un.unidecode = lambda s: s
    inf = types.ModuleType('inflect')
    inf.engine = lambda: types.SimpleNamespace(number_to_words=lambda *a, **k: '')
    plot = types.ModuleType('utils.plot')
    plot.plot_alignment = lambda *a, **k: None
    sys.modules.update({'jamo': jm, 'jamo.jamo': jmj, 'unidecode': un, 'inflect': inf, 'utils.plot': plot})
    for m in ('utils', 'utils.infolog', 'text', 'text.symbols', 'hparams'):        # the light stand-ins mtf installed: use the real packages now
        sys.modules.pop(m, None)
    sys.path.insert(0, REF)
    # TensorFlow surface of synthesizer.py
    tf.logging = types.SimpleNamespace(set_verbosity=lambda *_: None, ERROR=0)
    tf.ConfigProto = lambda **k: types.SimpleNamespace(gpu_options=types.SimpleNamespace())
    tf.Session = Session
    tf.global_variables_initializer = lambda: None
    tf.reset_default_graph = lambda: None
    tf.train.Saver = lambda *a, **k: types.SimpleNamespace(restore=lambda sess, path: None)
    named = contrib.placeholder

    def placeholder(dtype, shape=None, name=None):
        if name in ('inputs', 'input_lengths'):
            p = ArrayPlaceholder(dtype, [1, 2] if name == 'inputs' else [1], name)
            if name == 'input_lengths':
                p.value[:] = 2
            return p
        return named(dtype, shape, name)
    tf.placeholder = placeholder

    def placeholder_with_default(default, shape, name=None):
        p = ArrayPlaceholder(np.int32, [1], name)
        p.value = np.asarray(default).astype(np.int32)
        return p
    tf.placeholder_with_default = placeholder_with_default
    import synthesizer as ref                      # the reference module, unmodified
    assert ref.__file__.startswith(REF)

    class _Numpy(object):            # numpy < 1.24 built an object array from ragged input (synthesizer.py:94); newer numpy raises
        def __getattr__(self, k):
            return getattr(np, k)

        @staticmethod
        def array(x, *a, **k):
            try:
                return np.array(x, *a, **k)
            except ValueError:
                out = np.empty(len(x), dtype=object)
                for i, e in enumerate(x):
                    out[i] = e
                return out
    ref.np = _Numpy()
    ref.inv_spectrogram_tensorflow = lambda *a, **k: None            # TF Griffin-Lim: built at load, never fetched
    ref.inv_linear_spectrogram = lambda S, hp: np.zeros(S.shape[1] * 4, np.float32)       # Griffin-Lim of the trimmed spectrogram
    STATE['orig_initialize'] = ref.create_model(types.SimpleNamespace()).__class__.initialize

    def initialize(self, *a, **k):
        STATE['init'] = (self, a, k)
        out = STATE['orig_initialize'](self, *a, **k)
        STATE['names'] = {id(getattr(self, n)): n for n in ('linear_outputs', 'alignments', 'mel_outputs')}
        return out
    ref.create_model(types.SimpleNamespace()).__class__.initialize = initialize

    hp = synth.taco_tiny()
    ns = 2
    w = synth.make_taco_weights(hp, ns, seed=4321)
    tf.reset()
    tf.set_initial_values(w)
    tf.S.resolver = mtf.resolver
    contrib.FEED.clear()
    with tempfile.TemporaryDirectory() as d:
        with open(os.path.join(d, 'params.json'), 'w', encoding='euc-kr') as f:
            json.dump(dict(hp, max_iters=14), f)
        open(os.path.join(d, 'model.ckpt-700.data-00000-of-00001'), 'w').close()
        paths = [os.path.join(d, 'out', 's.wav')] * len(TEXTS)
        os.makedirs(os.path.join(d, 'out'))
        s = ref.Synthesizer()
        with contextlib.redirect_stdout(io.StringIO()):
            s.load(d, num_speakers=ns, checkpoint_step=None)
            s.synthesize(texts=TEXTS, paths=paths, speaker_ids=SPEAKERS, attention_trim=True, isKorean=True)
        out = {}
        files = sorted(glob.glob(os.path.join(d, 'out', '*.npy')))
        out['files'] = np.array([os.path.basename(p) for p in files])
        for i, p in enumerate(files):
            out['mel%d' % i] = np.load(p)
        model = STATE['init'][0]
        out['sequences'] = np.asarray(STATE_FEED['inputs'])
        out['input_lengths'] = np.asarray(STATE_FEED['input_lengths'])
        out['alignments'] = np.array(model.alignments)
        out['mel_outputs'] = np.array(model.mel_outputs)
    np.savez_compressed(os.path.join(HERE, 'ref_synth_main.npz'), texts=np.array(TEXTS), speakers=np.asarray(SPEAKERS), **out)
    print({k: np.shape(v) for k, v in out.items()}, out['files'])


if __name__ == '__main__':
    main()
