"""The WaveNet / audio hyper-parameters generate.py reads (reference hparams.py:18-34,57-79), as a
plain attribute bag.  The reference uses one global tf.contrib.training.HParams mutated in place by
load_hparams (utils/__init__.py:156-172); the same semantics are kept: `hparams` is a module-level
singleton and `load_hparams` overrides known keys from <checkpoint_dir>/params.json."""
import json
import os
import re


class HParams(object):
    def __init__(self, **kw):
        self.__dict__.update(kw)

    def values(self):
        return dict(self.__dict__)


hparams = HParams(
    name="Tacotron-Wavenet-Vocoder",
    # audio (hparams.py:18-34)
    sample_rate=24000, hop_size=300, fft_size=2048, win_size=1200, num_mels=80,
    preemphasize=True, preemphasis=0.97, min_level_db=-100, ref_level_db=20,
    signal_normalization=True, allow_clipping_in_normalization=True, symmetric_mels=True, max_abs_value=4.,
    rescaling=True, rescaling_max=0.999,
    # wavenet (hparams.py:57-79)
    filter_width=2, gc_channels=32, input_type="raw", scalar_input=True,
    dilations=[1, 2, 4, 8, 16, 32, 64, 128, 256, 512] * 5,
    residual_channels=32, dilation_channels=32, quantization_channels=256, out_channels=30,
    skip_channels=512, use_biases=True, initial_filter_width=32, upsample_factor=[5, 5, 12],
    # wavenet training (hparams.py:54-55,84-94)
    l2_regularization_strength=0, sample_size=15000, wavenet_batch_size=8, wavenet_learning_rate=1e-3, wavenet_decay_rate=0.5,
    wavenet_decay_steps=300000, wavenet_clip_gradients=False, optimizer='adam', num_steps=200000, skip_path_filter=False,
    max_checkpoints=3,
    # tacotron (hparams.py:124-166)
    cleaners='korean_cleaners', model_type='deepvoice', speaker_embedding_size=16, embedding_size=256, dropout_prob=0.5,
    enc_prenet_sizes=[256, 128], enc_bank_size=16, enc_bank_channel_size=128, enc_maxpool_width=2, enc_highway_depth=4,
    enc_rnn_size=128, enc_proj_sizes=[128, 128], enc_proj_width=3,
    attention_type='bah_mon_norm', attention_size=256, attention_state_size=256,
    dec_layer_num=2, dec_rnn_size=256, dec_prenet_sizes=[256, 128],
    post_bank_size=8, post_bank_channel_size=128, post_maxpool_width=2, post_highway_depth=4, post_rnn_size=128,
    post_proj_sizes=[256, 80], post_proj_width=3, reduction_factor=5, min_tokens=30, min_iters=30, max_iters=200,
    num_freq=1025, num_symbols=80,
)

PARAMS_NAME = "params.json"


def load_json(path, encoding='euc-kr'):
    # utils/__init__.py:173-185: tolerates trailing commas, euc-kr encoded
    with open(path, encoding=encoding) as f:
        content = f.read()
    content = re.sub(r",\s*}", "}", content)
    content = re.sub(r",\s*]", "]", content)
    return json.loads(content)


def load_hparams(hp, load_path, skip_list=()):
    """utils/__init__.py:156-172: update the singleton from <load_path>/params.json."""
    new = load_json(os.path.join(load_path, PARAMS_NAME))
    for key, value in new.items():
        if key in skip_list or key not in vars(hp):
            print("Skip {} because it not exists".format(key))
            continue
        if getattr(hp, key) != value:
            print("UPDATE {}: {} -> {}".format(key, getattr(hp, key), value))
            setattr(hp, key, value)
    return hp
