# coding: utf-8
"""train_vocoder.py: host logic on CPU (crop feeder = datasets/datafeeder_wavenet.py semantics, directory rules, checkpoint
naming) and, on the GPU, the whole loop train -> TF-format checkpoint -> restore -> generate."""
import os

import numpy as np
import pytest

from tacotron_wavenet_vocoder_korean_b200 import synth, tf_bundle, train_vocoder as tv
from tacotron_wavenet_vocoder_korean_b200.hparams import HParams


def _make_dirs(tmp_path, hop=6, num_mels=20, lengths=((40, 55, 70), (65, 45))):
    rs = np.random.RandomState(0)
    dirs = []
    for si, ls in enumerate(lengths):
        d = tmp_path / ('spk%d' % si)
        d.mkdir()
        lines = []
        for j, frames in enumerate(ls):
            mel = rs.randn(frames, num_mels).astype(np.float32)
            wav = (np.full(frames * hop, si + 0.01 * j)).astype(np.float32)
            name = 'utt-%d.npz' % j
            np.savez(str(d / name), audio=wav, mel=mel, time_steps=frames * hop)
            lines.append('a|b|c|%d|e|f|%s' % (frames * hop, name))
        (d / 'train.txt').write_text('\n'.join(lines) + '\n', encoding='utf-8')
        dirs.append(str(d))
    return dirs


def test_ensure_divisible_and_directory_rules():
    assert tv.ensure_divisible(15000, 300, True) == 15000 and tv.ensure_divisible(7680, 300, True) == 7500
    assert tv.ensure_divisible(7680, 300, False) == 7800
    with pytest.raises(ValueError):
        tv.validate_directories(tv.get_arguments(['--logdir', 'a', '--logdir_root', 'b']))
    with pytest.raises(ValueError):
        tv.validate_directories(tv.get_arguments(['--logdir', 'a', '--restore_from', 'b']))
    d = tv.validate_directories(tv.get_arguments(['--logdir_root', 'root']))
    assert d['logdir'].startswith(os.path.join('root', 'train')) and d['restore_from'] == d['logdir']
    d = tv.validate_directories(tv.get_arguments(['--logdir', 'keep']))
    assert d['logdir'] == d['restore_from'] == 'keep'


def test_crop_feeder_semantics(tmp_path):
    dirs = _make_dirs(tmp_path)
    hp = HParams(hop_size=6, sample_size=250, skip_path_filter=False)       # -> 246 samples = 41 frames
    f = tv.WavenetCropFeeder(dirs, batch_size=4, receptive_field=22, hparams=hp, gc_enable=True)
    assert f.sample_size == 246 and f.max_frames == 41
    # train.txt filter: only utterances longer than max(sample_size, rf) samples survive (datafeeder_wavenet.py:27)
    assert f.path_dict[dirs[0]] == ['utt-1.npz', 'utt-2.npz'] and f.path_dict[dirs[1]] == ['utt-0.npz', 'utt-1.npz']
    batches = f.make_batches()
    assert len(batches) == 32 and all(len(b) == 4 for b in batches)         # n*32/len(dirs) per speaker, batches of n
    spk = [e[2] for b in batches for e in b]
    assert spk.count(0) == spk.count(1) == 64
    wav, mel, gc = next(f)
    assert wav.shape == (4, 246) and mel.shape == (4, 41, 20) and gc.shape == (4,) and gc.dtype == np.int32
    for i in range(4):                                                       # the crop is hop-aligned and belongs to its speaker
        assert abs(wav[i, 0] - gc[i]) < 0.05 and np.all(wav[i] == wav[i, 0])
    # deterministic in the seed
    g1, g2_ = tv.WavenetCropFeeder(dirs, 4, 22, hp, gc_enable=True), tv.WavenetCropFeeder(dirs, 4, 22, hp, gc_enable=True)
    a, b = next(g1), next(g2_)
    assert all(np.array_equal(x, y) for x, y in zip(a, b))
    c = next(tv.WavenetCropFeeder(dirs, 4, 22, hp, gc_enable=True, seed=7))
    assert not all(np.array_equal(x, y) for x, y in zip(a, c))
    # single speaker: no global condition; too long a crop is an error, not an endless loop
    h = tv.WavenetCropFeeder(dirs[:1], 2, 22, hp, gc_enable=False)
    assert next(h)[2] is None
    with pytest.raises(ValueError):
        tv.WavenetCropFeeder(dirs, 4, 22, HParams(hop_size=6, sample_size=6000, skip_path_filter=False), gc_enable=True)


class _FakeTrainer(object):
    def __init__(self, names):
        self.variable_names = names
        self.global_step = 7
        self.s = {w: {n: np.full((2, 3), i + 10 * j, np.float32) for i, n in enumerate(names)} for j, w in enumerate(('params', 'ema', 'adam_m', 'adam_v'))}

    def state_dict(self, which='params'):
        return {k: v.copy() for k, v in self.s[which].items()}

    def load_state_dict(self, state, which='params', init_ema=True):
        self.s[which] = {k: np.array(v) for k, v in state.items()}


def test_checkpoint_layout_is_what_tf_saver_writes(tmp_path):
    names = ['wavenet/conv1d/kernel', 'wavenet/dilated_stack/layer0/dilation_layer/dense/bias']
    t = _FakeTrainer(names)
    prefix = tv.save(t, str(tmp_path), 7, {'sample_rate': 24000})
    r = tf_bundle.BundleReader(prefix)
    assert set(r.entries) == {n + s for n in names for s in ('', '/ExponentialMovingAverage')} | \
        {'optimizer/%s/%s' % (n, s) for n in names for s in ('Adam', 'Adam_1')} | {'optimizer/beta1_power', 'optimizer/beta2_power', 'global_step'}
    assert int(r.get_tensor('global_step')) == 7 and abs(float(r.get_tensor('optimizer/beta1_power')) - 0.9 ** 8) < 1e-7
    assert tf_bundle.checkpoint_state(str(tmp_path)) == prefix and os.path.exists(str(tmp_path / 'params.json'))
    # generate.py's loader sees only the raw variables; restore() brings back everything
    assert set(tf_bundle.load_variables(prefix)) == set(names)
    t2 = _FakeTrainer(names)
    t2.s = {w: {n: np.zeros((2, 3), np.float32) for n in names} for w in t2.s}
    assert tv.restore(t2, str(tmp_path)) == 7 and t2.global_step == 7
    for w in ('params', 'ema', 'adam_m', 'adam_v'):
        assert all(np.array_equal(t2.s[w][n], t.s[w][n]) for n in names)
    assert tv.restore(t2, str(tmp_path / 'nothing')) is None


@pytest.mark.gpu
def test_train_checkpoint_restore_generate_loop(tmp_path, monkeypatch):
    """3 + 2 training steps of a tiny model through the CLI, then the checkpoint drives the generation kernel."""
    import torch
    from tacotron_wavenet_vocoder_korean_b200 import hparams as hpmod
    from tacotron_wavenet_vocoder_korean_b200.generate import load_checkpoint
    from tacotron_wavenet_vocoder_korean_b200.wavenet import WaveNetModel
    dirs = _make_dirs(tmp_path, hop=6, num_mels=24, lengths=((60, 75, 90), (85, 65)))
    kw = synth.tiny_train(4)
    for k, v in dict(hop_size=6, sample_size=246, num_mels=24, dilations=kw['dilations'], residual_channels=16, dilation_channels=32,
                     skip_channels=64, initial_filter_width=8, gc_channels=8, upsample_factor=[2, 3], wavenet_batch_size=4,
                     num_steps=3, l2_regularization_strength=0).items():
        monkeypatch.setattr(hpmod.hparams, k, v)
    root = str(tmp_path / 'log')
    assert tv.main(['--data_dir', ','.join(dirs), '--logdir', root, '--checkpoint_every', '3', '--dtype', 'fp32']) == 3
    assert os.path.exists(os.path.join(root, 'model.ckpt-3.index'))
    monkeypatch.setattr(hpmod.hparams, 'num_steps', 5)
    assert tv.main(['--data_dir', ','.join(dirs), '--logdir', root, '--checkpoint_every', '5', '--dtype', 'fp32']) == 5      # resumes at 3
    state = load_checkpoint(root)                                                       # newest = step 5, raw variables only
    gkw = dict(kw, batch_size=2, global_condition_cardinality=2)
    assert set(state) == set(synth.weight_shapes(**gkw))
    net = WaveNetModel(train_mode=False, **gkw)
    net.load_state_dict(state)
    rs = np.random.RandomState(0)
    mel = rs.randn(2, 5, 24).astype(np.float32)
    lc = net.create_upsample(mel)
    out = net.generate(30, (2 * rs.rand(2, 1) - 1).astype(np.float32), rs.uniform(1e-5, 1 - 1e-5, (2, 30, 11)).astype(np.float32), lc_up=lc, gc_ids=[0, 1])
    assert torch.isfinite(out).all() and out.shape == (2, 30)
