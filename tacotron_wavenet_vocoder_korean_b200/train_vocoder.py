# coding: utf-8
"""`train_vocoder.py` of the reference (train_vocoder.py:26-186) on the B200 training step.

    python train_vocoder.py --data_dir ./data/moon,./data/son --logdir_root ./logdir-wavenet [--restore_from DIR]
                            [--checkpoint_every 1000] [--num_steps N] [--dtype bf16]

Kept from the reference: one speaker per data directory (speaker id = position in --data_dir, train_vocoder.py:82-88),
global conditioning iff more than one directory, the crop semantics of `datasets/datafeeder_wavenet.py` (hop-aligned
sample_size, one random window per npz, 32 batches' worth of examples per refill, shuffled), `net.add_loss` /
`net.add_optimizer` / `sess.run([global_step, loss, optimize])` (:122-123,169) -> WaveNetTrainer.train_step, the log line
(:172), a checkpoint every --checkpoint_every steps -- written as a TensorFlow checkpoint-V2 bundle
(`model.ckpt-<step>.{index,data-00000-of-00001}` + `checkpoint` + `params.json`, tf_bundle.py), variables, EMA shadows and Adam
slots under the names tf.train.Saver would use, so the reference (or generate.py here) restores it -- and the stop at
hparams.num_steps (:179).  Under torch.distributed (torchrun) every rank feeds its own crops and the gradients are averaged
with one all-reduce per step (the reference has no data parallelism).

Not kept: TF queues / threads (crops are cut on the host between steps; a 30 ms step leaves no feeder bottleneck at these
sizes), TensorBoard summaries, the timeline trace (`store_metadata`).
"""
import argparse
import json
import os
import sys
import time
from datetime import datetime
from glob import glob

import numpy as np


def ensure_divisible(length, divisible_by=256, lower=True):
    """datasets/datafeeder_wavenet.py:41-47."""
    if length % divisible_by == 0:
        return length
    if lower:
        return length - length % divisible_by
    return length + (divisible_by - length % divisible_by)


def get_path_dict(data_dirs, min_length, skip_path_filter=False):
    """datasets/datafeeder_wavenet.py:15-36: npz files per speaker directory; with train.txt present only utterances longer
    than `min_length` samples (column 3 = time_steps, column 6 = npz name) are kept."""
    out = {}
    for d in data_dirs:
        meta = os.path.join(d, 'train.txt')
        if not skip_path_filter and os.path.exists(meta):
            names = []
            with open(meta, encoding='utf-8') as f:
                for line in f:
                    parts = line.strip().split('|')
                    if len(parts) > 6 and int(parts[3]) > min_length:
                        names.append(parts[6])
            out[d] = names
        else:
            out[d] = sorted(os.path.basename(p) for p in glob(os.path.join(d, '*.npz')))
    return out


class WavenetCropFeeder(object):
    """Host-side restatement of DataFeederWavenet (datasets/datafeeder_wavenet.py:50-176) as an iterator of
    (wav (N, sample_size) float32, mel (N, sample_size / hop, num_mels) float32, speaker ids (N,) int32 or None)."""

    def __init__(self, data_dirs, batch_size, receptive_field, hparams, gc_enable=False, seed=123, crop_rng=None):
        self.data_dirs = list(data_dirs)
        self.batch_size = batch_size
        self.hop_size = hparams.hop_size
        self.sample_size = ensure_divisible(hparams.sample_size, self.hop_size, True)          # :58
        self.max_frames = self.sample_size // self.hop_size
        self.gc_enable = gc_enable
        self.skip_path_filter = getattr(hparams, 'skip_path_filter', False)
        self.rng = np.random.RandomState(seed)                                                # :65
        self.crop_rng = crop_rng if crop_rng is not None else self.rng                        # :153 crops with the GLOBAL numpy RNG
        self._offset = {d: 2 for d in self.data_dirs}                                         # :66 defaultdict(lambda: 2)
        self.data_dir_to_id = {d: i for i, d in enumerate(self.data_dirs)}
        self.path_dict = get_path_dict(self.data_dirs, max(self.sample_size, receptive_field), self.skip_path_filter)
        for d in self.data_dirs:
            if not self.path_dict[d]:
                raise ValueError("no usable npz files in %s" % d)
        self._pending = []

    def _get_next_example(self, data_dir):
        paths = self.path_dict[data_dir]
        tried = 0
        while True:
            if self._offset[data_dir] >= len(paths):
                self._offset[data_dir] = 0
                self.rng.shuffle(paths)
            path = os.path.join(data_dir, paths[self._offset[data_dir]])
            self._offset[data_dir] += 1
            tried += 1
            if tried > 4 * len(paths) + 8:
                raise ValueError("no npz in %s is longer than sample_size %d" % (data_dir, self.sample_size))
            if not os.path.exists(path):
                continue
            data = np.load(path)
            mel = data['mel']
            if len(mel) >= self.max_frames and (not self.skip_path_filter or 'time_steps' not in data.files
                                                or int(data['time_steps']) > self.sample_size):
                break
        wav = np.asarray(data['audio'], np.float32).reshape(-1)
        assert len(wav) % len(mel) == 0 and len(wav) // len(mel) == self.hop_size, "audio / mel not hop-aligned (datafeeder_wavenet.py:38)"
        s = self.crop_rng.randint(0, len(mel) - self.max_frames + 1)                          # :153
        ts = s * self.hop_size
        ex = (wav[ts:ts + self.hop_size * self.max_frames], np.asarray(mel[s:s + self.max_frames], np.float32))
        return ex + ((self.data_dir_to_id[data_dir],) if self.gc_enable else ())

    def make_batches(self):
        """:115-128: 32 batches' worth of examples, the same share from every speaker, shuffled; partial batches are dropped
        (the reference would feed them to a fixed-batch graph and fail)."""
        n = self.batch_size
        examples = []
        for d in self.data_dirs:
            examples.extend(self._get_next_example(d) for _ in range(int(n * 32 // len(self.data_dirs))))
        self.rng.shuffle(examples)
        return [examples[i:i + n] for i in range(0, len(examples) - n + 1, n)]

    def __iter__(self):
        return self

    def __next__(self):
        if not self._pending:
            self._pending = self.make_batches()
        batch = self._pending.pop(0)
        wav = np.stack([b[0] for b in batch])
        mel = np.stack([b[1] for b in batch])
        gc = np.array([b[2] for b in batch], np.int32) if self.gc_enable else None
        return wav, mel, gc


def checkpoint_tensors(trainer):
    """Everything tf.train.Saver(var_list=tf.global_variables()) stores (train_vocoder.py:131): the variables, their EMA
    shadows (wavenet/model.py:346), the Adam slots and counters (model.py:325), global_step."""
    out = {}
    p, ema, m, v = (trainer.state_dict(w) for w in ('params', 'ema', 'adam_m', 'adam_v'))
    for name in p:
        out[name] = p[name]
        out[name + '/ExponentialMovingAverage'] = ema[name]
        out['optimizer/' + name + '/Adam'] = m[name]
        out['optimizer/' + name + '/Adam_1'] = v[name]
    t = trainer.global_step
    ta = getattr(trainer, 'adam_t', t)          # Adam's own step count: survives a global_step reset (--restore_from + new logdir)
    out['optimizer/beta1_power'] = np.array(0.9 ** (ta + 1), np.float32)
    out['optimizer/beta2_power'] = np.array(0.999 ** (ta + 1), np.float32)
    out['global_step'] = np.array(t, np.int32)
    return out


def save(trainer, logdir, step, hparams_dict=None):
    """utils/__init__.py:62-72 `save(saver, sess, logdir, step)` + the `checkpoint` state file + params.json."""
    from . import tf_bundle
    os.makedirs(logdir, exist_ok=True)
    prefix = os.path.join(logdir, 'model.ckpt-%d' % step)
    print('Storing checkpoint to {} ...'.format(logdir), end="")
    sys.stdout.flush()
    tf_bundle.write_bundle(prefix, checkpoint_tensors(trainer))
    # tf.train.Saver(max_to_keep=hparams.max_checkpoints) (train_vocoder.py:126): keep the newest few, list them all
    keep_n = int((hparams_dict or {}).get('max_checkpoints', 3) or 3)
    steps = sorted({int(f[len('model.ckpt-'):-len('.index')]) for f in os.listdir(logdir)
                    if f.startswith('model.ckpt-') and f.endswith('.index') and f[len('model.ckpt-'):-len('.index')].isdigit()})
    for old in steps[:-keep_n]:
        for f in os.listdir(logdir):
            if f.startswith('model.ckpt-%d.' % old):
                os.remove(os.path.join(logdir, f))
    kept = steps[-keep_n:]
    with open(os.path.join(logdir, 'checkpoint'), 'w') as f:
        f.write('model_checkpoint_path: "model.ckpt-%d"\n' % step)
        for k in kept:
            f.write('all_model_checkpoint_paths: "model.ckpt-%d"\n' % k)
    if hparams_dict is not None:
        with open(os.path.join(logdir, 'params.json'), 'w', encoding='utf-8') as f:
            json.dump(hparams_dict, f, indent=4, sort_keys=True, ensure_ascii=False)
    print(' Done.')
    return prefix


def restore(trainer, logdir):
    """utils/__init__.py:75-90 `load(saver, sess, logdir)`: newest checkpoint -> variables, EMA shadows, Adam slots, step."""
    from . import tf_bundle
    prefix = tf_bundle.checkpoint_state(logdir) if logdir else None
    if prefix is None or not os.path.exists(prefix + '.index'):
        print(" No checkpoint found.")
        return None
    print("  Checkpoint found: {}".format(prefix))
    r = tf_bundle.BundleReader(prefix)
    names = trainer.variable_names
    trainer.load_state_dict({n: r.get_tensor(n) for n in names})
    for which, fmt in (('ema', '%s/ExponentialMovingAverage'), ('adam_m', 'optimizer/%s/Adam'), ('adam_v', 'optimizer/%s/Adam_1')):
        if all(r.has_tensor(fmt % n) for n in names):
            trainer.load_state_dict({n: r.get_tensor(fmt % n) for n in names}, which=which)
    step = tf_bundle.global_step_of(prefix)
    print("  Global step was: {}".format(step))
    trainer.global_step = step
    # Adam's step count from the saved beta powers (the reference's Saver restores them independently of global_step)
    trainer.adam_t = step
    if r.has_tensor('optimizer/beta1_power'):
        b1p = float(np.asarray(r.get_tensor('optimizer/beta1_power')).reshape(-1)[0])
        if 0.0 < b1p < 1.0:
            trainer.adam_t = max(int(round(np.log(b1p) / np.log(0.9))) - 1, 0)
    return step


def get_arguments(argv=None):
    parser = argparse.ArgumentParser(description='WaveNet example network')
    parser.add_argument('--data_dir', type=str, default='./data/moon,./data/son', help='The directories containing the preprocessed npz files, one per speaker.')
    parser.add_argument('--logdir', type=str, default=None)
    parser.add_argument('--logdir_root', type=str, default=None)
    parser.add_argument('--restore_from', type=str, default=None)
    parser.add_argument('--checkpoint_every', type=int, default=1000)
    parser.add_argument('--num_steps', type=int, default=None, help='overrides hparams.num_steps')
    parser.add_argument('--dtype', type=str, default='bf16', choices=['bf16', 'fp32'])
    return parser.parse_args(argv)


def validate_directories(config):
    """utils/__init__.py:100-140: --logdir excludes --logdir_root / --restore_from; default root ./logdir-wavenet;
    a new run goes to <root>/train/<timestamp> and restores from --restore_from (or from itself).  Difference kept on purpose:
    the reference writes params.json when the run directory is created and re-reads it when --logdir continues a run; here
    params.json is written next to every checkpoint (`save`) and a continued run keeps the hparams of the current process."""
    if config.logdir and config.logdir_root:
        raise ValueError("--logdir and --logdir_root cannot be specified at the same time.")
    if config.logdir and config.restore_from:
        raise ValueError("--logdir continues a run in place, so it cannot be combined with --restore_from (that would overwrite the "
                         "restored model); use --logdir_root to start a new dated run from --restore_from.")
    root = config.logdir_root or './logdir-wavenet'
    logdir = config.logdir or os.path.join(root, 'train', "{0:%Y-%m-%dT%H-%M-%S}".format(datetime.now()))
    restore_from = config.restore_from or logdir
    return dict(logdir=logdir, logdir_root=config.logdir_root, restore_from=restore_from)


def main(argv=None):
    import torch
    import torch.distributed as dist
    from .hparams import hparams
    from .wavenet import WaveNetModel
    config = get_arguments(argv)
    config.data_dir = config.data_dir.split(",")
    try:
        directories = validate_directories(config)
    except ValueError as e:
        print("Some arguments are wrong:")
        print(str(e))
        return
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    if world > 1:
        torch.cuda.set_device(int(os.environ.get('LOCAL_RANK', '0')))
        dist.init_process_group('nccl')
    logdir, restore_from = directories['logdir'], directories['restore_from']
    is_overwritten_training = logdir != restore_from
    num_speakers = len(config.data_dir)
    gc_enable = num_speakers > 1
    rf = WaveNetModel.calculate_receptive_field(hparams.filter_width, hparams.dilations, hparams.scalar_input, hparams.initial_filter_width)
    reader = WavenetCropFeeder(config.data_dir, hparams.wavenet_batch_size, rf, hparams, gc_enable=gc_enable, seed=123 + rank)
    net = WaveNetModel(batch_size=hparams.wavenet_batch_size, dilations=hparams.dilations, filter_width=hparams.filter_width,
                       residual_channels=hparams.residual_channels, dilation_channels=hparams.dilation_channels,
                       quantization_channels=hparams.quantization_channels, out_channels=hparams.out_channels,
                       skip_channels=hparams.skip_channels, use_biases=hparams.use_biases, scalar_input=hparams.scalar_input,
                       initial_filter_width=hparams.initial_filter_width, global_condition_channels=hparams.gc_channels if gc_enable else None,
                       global_condition_cardinality=num_speakers if gc_enable else None, local_condition_channels=hparams.num_mels,
                       upsample_factor=hparams.upsample_factor, train_mode=True)
    l2 = None if hparams.l2_regularization_strength == 0 else hparams.l2_regularization_strength       # train_vocoder.py:118-119
    trainer = net.trainer(reader.sample_size, config.dtype)
    start_step = restore(trainer, restore_from)
    if is_overwritten_training or start_step is None:
        trainer.global_step = 0                                                                        # :139-143
    trainer.sync_params(0)
    num_steps = config.num_steps if config.num_steps is not None else getattr(hparams, 'num_steps', 200000)
    step = trainer.global_step
    while True:
        start_time = time.time()
        wav, mel, gc = next(reader)
        loss = trainer.train_step(wav, mel, gc, hparams, l2)
        step = trainer.global_step
        loss_value = float(loss.item())
        if rank == 0:
            print('step {:d} - loss = {:.3f}, ({:.3f} sec/step)'.format(step, loss_value, time.time() - start_time))
            if step % config.checkpoint_every == 0:
                save(trainer, logdir, step, hparams.values())
        if step >= num_steps:
            break
    if world > 1:
        dist.destroy_process_group()
    return step


if __name__ == '__main__':
    main()
    print('Done')
