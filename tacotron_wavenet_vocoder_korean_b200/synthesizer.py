# coding: utf-8
"""Drop-in for the reference's synthesizer.py: `Synthesizer.load / synthesize / close` (synthesizer.py:30-200)
and the CLI (synthesizer.py:370-388), with the Tacotron graph on sm_100a kernels (libtaco_b200.so).

    python synthesizer.py --load_path logdir-tacotron/moon+son --num_speakers 2 --speaker_id 0 --text "..."

What is kept: text -> ids (text_to_sequence + _prepare_inputs), input_lengths = argmax(ids == EOS) + 1
(synthesizer.py:126), ONE evaluation of [linear_outputs, alignments, mel_outputs] (:129-160), the attention-based
end trim (:235-256) and `np.save(mel)` next to the output (:279-280) -- the `.npy` generate.py takes as --mel.
What is not: Griffin-Lim (`inv_linear_spectrogram`, a separate vocoder), alignment
plots (matplotlib) and the manual-attention post-passes, which crash in the reference;
`base_alignment_path` (fed manual alignments) is supported.  The checkpoint is the reference's own
`model.ckpt-<step>.{index,data-00000-of-00001}` (read by tf_bundle.py, no TensorFlow) or `weights[-<step>].npz` keyed by
the TF variable names, next to `params.json` in `load_path`.
"""
import argparse
import datetime
import os

import numpy as np

from .hparams import hparams, load_hparams, PARAMS_NAME
from .tacotron import create_model, get_most_recent_checkpoint, load_weights
from .text import text_to_sequence, prepare_inputs


def get_time():
    return datetime.datetime.now().strftime("%Y-%m-%d_%H-%M-%S")


def attention_trim_index(alignment, seq_len, reduction_factor):
    """End-of-sentence trimming of synthesizer.py:235-256 for one alignment (T_in, T_dec): the number of spectrogram
    frames to keep.  With a = per-decoder-step argmax and `last` = min(seq_len - 1, max(a)), decoding is considered finished
    at the first step j (before the final one) where either `last` has been attended min(#steps on `last`, 5) times, or
    `last` is attended and the next step moves past it; otherwise at the final step.  Frames kept = r * j + 3."""
    a = np.asarray(alignment).argmax(0)
    if a.size <= 1:
        return 3
    last = min(int(seq_len) - 1, int(a.max()))
    on_last = a[:-1] == last
    held = min(int((a == last).sum()), 5)
    stop = (np.cumsum(on_last) >= held) | (on_last & (a[1:] > last))
    j = int(np.argmax(stop)) if stop.any() else a.size - 1
    return reduction_factor * j + 3


class Synthesizer(object):
    def __init__(self):
        self.model = None
        self.num_speakers = None

    def close(self):
        self.model = None

    def load(self, checkpoint_path, num_speakers=2, checkpoint_step=None, model_name='tacotron', weights=None):
        """synthesizer.py:34-70.  `weights` (a state dict) bypasses the checkpoint file (tests, benchmarks)."""
        self.num_speakers = num_speakers
        if weights is None:
            if os.path.isdir(checkpoint_path):
                load_path = checkpoint_path
                checkpoint_path = get_most_recent_checkpoint(load_path, checkpoint_step)
            else:
                load_path = os.path.dirname(checkpoint_path)
            if os.path.exists(os.path.join(load_path, PARAMS_NAME)):
                load_hparams(hparams, load_path)
            print('Loading checkpoint: %s' % checkpoint_path)
            weights = load_weights(checkpoint_path)
        print('Constructing model: %s' % model_name)
        self.model = create_model(hparams)
        self.model.load_state_dict(weights)

    def synthesize(self, texts=None, tokens=None, base_path=None, paths=None, speaker_ids=None, start_of_sentence=None,
                   end_of_sentence=True, pre_word_num=0, post_word_num=0, pre_surplus_idx=0, post_surplus_idx=1,
                   use_short_concat=False, manual_attention_mode=0, base_alignment_path=None, librosa_trim=False,
                   attention_trim=True, isKorean=True):
        if use_short_concat or manual_attention_mode or librosa_trim:
            raise NotImplementedError("use_short_concat / manual_attention_mode / librosa_trim are not built "
                                      "(they crash in the reference, SURVEY.md App. E-6)")
        if isinstance(texts, str):
            texts = [texts]
        if texts is not None and tokens is None:
            sequences = prepare_inputs([text_to_sequence(t) for t in texts])
        elif tokens is not None:
            sequences = np.asarray(tokens)
        else:
            raise ValueError("texts or tokens required")
        N = len(sequences)
        if paths is None:
            paths = [None] * N
        if texts is None:
            texts = [None] * N
        input_lengths = [int(np.argmax(a == 1)) + 1 for a in sequences]           # synthesizer.py:126
        m = self.model
        if base_alignment_path is None:
            m.is_manual_attention, m.manual_alignments = False, None
        else:
            alignment_path = os.path.join(os.path.basename(base_path), base_alignment_path)
            man = [np.load("{}{}.npy".format(alignment_path, idx)) for idx in range(N)]
            m.is_manual_attention, m.manual_alignments = True, np.transpose(man, [0, 2, 1])
        if isinstance(speaker_ids, dict):
            raise NotImplementedError("speaker mixing through a dict of speaker ids (synthesizer.py:151-156) crashes in the reference")
        spk = None if speaker_ids is None else np.asarray(speaker_ids, np.int32)
        m.initialize(sequences, input_lengths, self.num_speakers, spk, rnn_decoder_test_mode=True)
        linears = m.linear_outputs.cpu().numpy()
        alignments = m.alignments.cpu().numpy()
        mels = m.mel_outputs.cpu().numpy()
        r = hparams.reduction_factor
        results = []
        for idx in range(N):
            lin, al, mel = linears[idx], alignments[idx], mels[idx]
            if attention_trim and end_of_sentence:
                end = attention_trim_index(al, len(sequences[idx]), r)
                lin, mel = lin[:end], mel[:end]
            out = {'mel': mel, 'linear': lin, 'alignment': al, 'text': texts[idx]}
            if paths[idx] or base_path:
                if paths[idx]:
                    root, ext = os.path.splitext(paths[idx])
                    current = "%s.%d%s" % (root, idx, ext or '.wav')
                else:
                    os.makedirs(base_path, exist_ok=True)
                    # the reference names this "<base_path>/<time>.wav" for every sentence of the batch (synthesizer.py:206,265), so
                    # sentences synthesised within one second overwrite each other; the sentence index keeps them apart here
                    current = "{}/{}.{}.wav".format(base_path, get_time(), idx)
                mel_path = current.replace(".wav", ".npy")
                np.save(mel_path, mel)                                            # synthesizer.py:279-280
                out['mel_path'] = mel_path
            results.append(out)
        return results


def main(argv=None):
    parser = argparse.ArgumentParser()
    parser.add_argument('--load_path', required=True)
    parser.add_argument('--sample_path', default="logdir-tacotron/generate")
    parser.add_argument('--text', required=True)
    parser.add_argument('--num_speakers', default=1, type=int)
    parser.add_argument('--speaker_id', default=0, type=int)
    parser.add_argument('--checkpoint_step', default=None, type=int)
    parser.add_argument('--is_korean', default=True, type=lambda s: str(s).lower() in ('1', 'true', 'yes', 'y'))
    parser.add_argument('--base_alignment_path', default=None)
    config = parser.parse_args(argv)
    os.makedirs(config.sample_path, exist_ok=True)
    synthesizer = Synthesizer()
    synthesizer.load(config.load_path, config.num_speakers, config.checkpoint_step)
    res = synthesizer.synthesize(texts=[config.text], base_path=config.sample_path, speaker_ids=[config.speaker_id],
                                 attention_trim=True, base_alignment_path=config.base_alignment_path, isKorean=config.is_korean)[0]
    print("mel %s -> %s" % (res['mel'].shape, res.get('mel_path')))


if __name__ == "__main__":
    main()
