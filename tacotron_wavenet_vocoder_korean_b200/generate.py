#  coding: utf-8
"""WaveNet generation script -- the reference's generate.py entry point on the B200 path.

Same CLI (generate.py:38-79): positional checkpoint_dir, --temperature, --logdir, --wav_out_path,
--batch_size, --wav_seed, --mel, --gc_cardinality, --gc_id; writes logdir/generate/<timestamp>/test-{i}.wav
(generate.py:109,261).  The per-sample Python loop around sess.run (generate.py:202-233) is ONE launch of
the persistent kernel (WaveNetModel.generate).

Differences forced by the environment, stated plainly:
  * checkpoints: TensorFlow tensor-bundle files cannot be read here (no TF); checkpoint_dir must hold
    params.json (as the reference writes it) and `weights.npz` keyed by the TF variable names.
    `--synthetic_weights` fabricates seeded weights of the configured architecture (no trained
    checkpoint ships with the reference).
  * --wav_seed: a 16-bit PCM .wav at the model sample rate (librosa resampling / silence trimming of
    generate.py:88-103 are not reproduced).
  * randomness: the reference is unseeded; --seed makes the run reproducible.
"""
import argparse
import os
import time
from datetime import datetime

import numpy as np


def get_arguments(argv=None):
    def _ensure_positive_float(f):
        if float(f) < 0:
            raise argparse.ArgumentTypeError('Argument must be greater than zero')
        return float(f)

    parser = argparse.ArgumentParser(description='WaveNet generation script')
    parser.add_argument('checkpoint_dir', type=str, help='Which model checkpoint to generate from')
    parser.add_argument('--temperature', type=_ensure_positive_float, default=1.0, help='Sampling temperature')
    parser.add_argument('--logdir', type=str, default='./logdir-wavenet', help='Directory in which to store the output.')
    parser.add_argument('--wav_out_path', type=str, default=None, help='Path to output wav file')
    parser.add_argument('--batch_size', type=int, default=1, help='batch size')
    parser.add_argument('--wav_seed', type=str, default=None, help='The wav file to start generation from')
    parser.add_argument('--mel', type=str, default=None, help='mel input')
    parser.add_argument('--gc_cardinality', type=int, default=None, help='Number of categories upon which we globally condition.')
    parser.add_argument('--gc_id', type=int, default=None, help='ID of category to generate, if globally conditioned.')
    parser.add_argument('--seed', type=int, default=None, help='RNG seed (the reference is unseeded)')
    parser.add_argument('--synthetic_weights', action='store_true', help='use seeded random weights instead of weights.npz')
    arguments = parser.parse_args(argv)
    return arguments


def create_seed(filename, sample_rate, quantization_channels, window_size, scalar_input):
    """generate.py:88-103 without librosa: first `window_size` samples of a PCM wav."""
    from scipy.io import wavfile
    sr, data = wavfile.read(filename)
    if sr != sample_rate:
        raise ValueError('seed wav is %d Hz, the model runs at %d Hz (no resampler here)' % (sr, sample_rate))
    if data.ndim > 1:
        data = data.mean(axis=1)
    audio = data.astype(np.float32) / 32768.0 if data.dtype.kind == 'i' else data.astype(np.float32)
    audio = audio[:window_size]
    if scalar_input:
        return audio
    from .wavenet import mu_law_encode
    import torch
    return mu_law_encode(torch.from_numpy(audio), quantization_channels).cpu().numpy().astype(np.float32)


def load_checkpoint(checkpoint_dir, synthetic=None):
    """utils/__init__.py:75-90 `load(saver, sess, logdir)`: the TF checkpoint named by `<dir>/checkpoint` (or the newest
    `model.ckpt-<step>.index`) is read with tf_bundle (no TensorFlow); `weights.npz` keyed by the same variable names
    is the fallback interchange file; `synthetic` (a callable) supplies seeded weights for benchmarks."""
    from . import tf_bundle
    if synthetic is not None:
        return synthetic()
    prefix = tf_bundle.checkpoint_state(checkpoint_dir)
    if prefix is not None and os.path.exists(prefix + '.index'):
        print("  Checkpoint found: {}".format(prefix))
        print("  Global step was: {}".format(tf_bundle.global_step_of(prefix)))
        return tf_bundle.load_variables(prefix)
    wpath = os.path.join(checkpoint_dir, 'weights.npz')
    if not os.path.exists(wpath):
        raise FileNotFoundError('no model.ckpt-<step>.index / weights.npz in %s; pass --synthetic_weights for seeded '
                                'random weights' % checkpoint_dir)
    print('Restoring model from {}'.format(checkpoint_dir))
    return dict(np.load(wpath))


def main(argv=None):
    import torch
    from .hparams import hparams, load_hparams
    from .wavenet import WaveNetModel, mu_law_decode
    from . import audio, synth

    config = get_arguments(argv)
    started_datestring = "{0:%Y-%m-%dT%H-%M-%S}".format(datetime.now())
    logdir = os.path.join(config.logdir, 'generate', started_datestring)
    os.makedirs(logdir, exist_ok=True)
    if os.path.exists(os.path.join(config.checkpoint_dir, 'params.json')):
        load_hparams(hparams, config.checkpoint_dir)
    if hparams.gc_channels is not None:
        if config.gc_cardinality is None:
            raise ValueError("Globally conditioning but gc_cardinality not specified. Use --gc_cardinality=377 for full VCTK corpus.")
        if config.gc_id is None:
            raise ValueError("Globally conditioning, but global condition was not specified. Use --gc_id to specify global condition.")
    if config.mel is None:
        raise ValueError("--mel is required: the incremental graph always consumes a local-condition row (generate.py:151)")

    scalar_input = hparams.scalar_input
    kwargs = dict(batch_size=config.batch_size, dilations=hparams.dilations, filter_width=hparams.filter_width,
                  residual_channels=hparams.residual_channels, dilation_channels=hparams.dilation_channels,
                  quantization_channels=hparams.quantization_channels, out_channels=hparams.out_channels,
                  skip_channels=hparams.skip_channels, use_biases=hparams.use_biases, scalar_input=hparams.scalar_input,
                  initial_filter_width=hparams.initial_filter_width, global_condition_channels=hparams.gc_channels,
                  global_condition_cardinality=config.gc_cardinality, local_condition_channels=hparams.num_mels,
                  upsample_factor=hparams.upsample_factor)
    net = WaveNetModel(train_mode=False, **kwargs)
    state = load_checkpoint(config.checkpoint_dir, None if not config.synthetic_weights else (lambda: synth.make_weights(**kwargs)))
    net.load_state_dict(state)
    net.queue_initializer()

    rs = np.random.RandomState(config.seed)
    N = net.batch_size
    mel_input = np.load(config.mel)
    sample_size = mel_input.shape[0] * hparams.hop_size
    mel_input = np.tile(mel_input, (N, 1, 1)).astype(np.float32)                    # generate.py:153
    # generate.py:155,200 materialise create_upsample(mel); here the kernel evaluates it per step from TMA-staged mel frames
    Q = hparams.quantization_channels
    if config.wav_seed:
        seed = create_seed(config.wav_seed, hparams.sample_rate, Q, net.receptive_field, scalar_input)
        seed = np.asarray(seed, np.float32)[-net.receptive_field:]
        forced = np.tile(seed[None, :], (N, 1))
        print('Priming generation...')
    else:
        # silence with a single random sample at the end; only the last element is ever fed (generate.py:184-192,204)
        if scalar_input:
            forced = (2 * rs.rand(N) - 1).reshape(N, 1).astype(np.float32)
        else:
            forced = rs.randint(Q, size=N).reshape(N, 1).astype(np.float32)
    n_prime = forced.shape[1] - 1
    T = n_prime + sample_size
    if scalar_input:
        uniforms = rs.uniform(1e-5, 1 - 1e-5, (N, T, hparams.out_channels // 3 + 1)).astype(np.float32)
    else:
        uniforms = rs.random_sample((N, T))
    gc = [config.gc_id] * N if hparams.gc_channels is not None else None
    start_time = time.time()
    out = net.generate(T, forced, uniforms, mel=mel_input, lc_shift=n_prime, gc_ids=gc, temperature=config.temperature)
    out = out[:, n_prime:]
    torch.cuda.synchronize()
    print('Generated {} samples x {} rows in {:.3f} sec'.format(sample_size, N, time.time() - start_time))

    if hparams.input_type == 'raw':
        wav = out
    elif hparams.input_type == 'mulaw':
        wav = mu_law_decode(out, Q, quantization=False)
    else:  # 'mulaw-quantize'
        wav = mu_law_decode(out, Q, quantization=True)
    wav = wav.cpu().numpy()
    for i in range(N):
        path = config.wav_out_path if (config.wav_out_path and N == 1) else logdir + '/test-{}.wav'.format(i)
        audio.save_wav(wav[i], path, hparams.sample_rate)
    print('Finished generating.')
    return wav


if __name__ == '__main__':
    s = time.time()
    main()
    print(time.time() - s, 'sec')
