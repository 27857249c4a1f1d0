# coding: utf-8
"""Generates tests/golden/ref_trim.npz by running THE REFERENCE'S OWN synthesizer.plot_graph_and_save_audio (synthesizer.py:198-277)
-- the attention-based end-of-sentence trimming that decides how many spectrogram / mel frames an utterance keeps -- on random
alignment paths.  Its collaborators (plotting, Griffin-Lim, wav writing, the model / text / feeder packages) are replaced by
recorders before `synthesizer` is imported; the function body itself runs unmodified.

    python tests/golden/make_reference_trim_golden.py        (build container only)
"""
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get('REFERENCE_ROOT', '/root/reference')
sys.path.insert(0, HERE)
import make_reference_audio_golden as mra      # noqa: E402
import tf_numpy_shim as tf                     # noqa: E402

T_DEC, R_FACTOR = 40, 5


def cases(seed=3, n=400):
    """Per-step argmax paths (T_DEC,) and sequence lengths: monotonic walks that reach the last token early / late / never,
    walks that overshoot seq_len - 1 (padding attended), constant paths, noisy non-monotonic ones."""
    rs = np.random.RandomState(seed)
    paths, lens = [], []
    for i in range(n):
        L = int(rs.randint(2, 25))
        kind = i % 5
        if kind == 0:
            p = np.minimum(np.cumsum(rs.rand(T_DEC) < rs.uniform(0.2, 0.9)), L - 1)
        elif kind == 1:
            p = np.minimum(np.cumsum(rs.rand(T_DEC) < 0.7), L + 3)             # runs past the sentence into the padding
        elif kind == 2:
            p = np.full(T_DEC, rs.randint(0, L + 2))
        elif kind == 3:
            p = np.clip(np.cumsum(rs.randint(-1, 3, T_DEC)), 0, L + 1)
        else:
            p = rs.randint(0, L + 2, T_DEC)
        paths.append(p.astype(np.int64))
        lens.append(L)
    return np.stack(paths), np.asarray(lens, np.int64)


def main():
    mra.install_stubs()
    tf.logging = types.SimpleNamespace(set_verbosity=lambda *_: None, ERROR=0)
    sys.path.insert(0, REF)
    kept = []
    stubs = {
        'utils.plot': dict(plot_alignment=lambda *a, **k: None),
        'utils.audio': dict(save_wav=lambda audio, path, sr: kept.append(len(audio)), inv_linear_spectrogram=lambda S, hp: np.zeros(S.shape[1]),
                            inv_preemphasis=None, inv_spectrogram_tensorflow=None),
        'tacotron': dict(create_model=None, get_most_recent_checkpoint=None),
        'text': dict(text_to_sequence=None, sequence_to_text=None),
        'text.korean': dict(tokenize=None),
        'datasets': dict(),
        'datasets.datafeeder_tacotron': dict(_prepare_inputs=None),
        'tqdm': dict(tqdm=lambda x, **k: x),
    }
    for name, attrs in stubs.items():
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        if name in ('text', 'datasets'):
            m.__path__ = []
        sys.modules[name] = m
    for m in ('utils', 'hparams', 'synthesizer'):
        sys.modules.pop(m, None)
    import synthesizer as ref                    # the reference module, unmodified
    assert ref.__file__.startswith(REF)
    ref.hparams.reduction_factor = R_FACTOR
    paths, lens = cases()
    out = []
    for p, L in zip(paths, lens):
        n_in = int(max(L, p.max() + 1))
        alignment = np.zeros((n_in, T_DEC), np.float32)
        alignment[p, np.arange(T_DEC)] = 1.0
        wav = np.zeros((T_DEC * R_FACTOR, 7), np.float32)          # the linear spectrogram (frames, bins)
        mel = np.zeros((T_DEC * R_FACTOR, 3), np.float32)
        kept.clear()
        ref.plot_graph_and_save_audio((0, (wav, alignment, None, None, np.zeros(L, np.int64), mel)), base_path=None,
                                      end_of_sentence=True, attention_trim=True)
        out.append(kept[0])
    np.savez_compressed(os.path.join(HERE, 'ref_trim.npz'), paths=paths, lens=lens, kept=np.asarray(out, np.int64),
                        t_dec=np.int64(T_DEC), reduction_factor=np.int64(R_FACTOR))
    print(np.bincount(np.asarray(out))[:60], min(out), max(out))


if __name__ == '__main__':
    main()
