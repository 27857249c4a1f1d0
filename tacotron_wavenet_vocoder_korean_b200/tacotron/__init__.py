# Same exports as the reference's tacotron/__init__.py:1-13.
import os
from glob import glob

from .tacotron import Tacotron


def create_model(hparams):
    return Tacotron(hparams)


def get_most_recent_checkpoint(checkpoint_dir, checkpoint_step=None):
    """tacotron/__init__.py:11-13, synthesizer.py:289-299: the newest `model.ckpt-<step>` TF checkpoint prefix (read by
    tf_bundle); without one, the interchange file weights-<step>.npz / weights.npz keyed by the TF variable names."""
    from .. import tf_bundle
    prefix = tf_bundle.get_most_recent_checkpoint(checkpoint_dir, checkpoint_step)
    if prefix is not None and os.path.exists(prefix + '.index'):
        return prefix
    if checkpoint_step is not None:
        return os.path.join(checkpoint_dir, "weights-%d.npz" % checkpoint_step)
    paths = glob(os.path.join(checkpoint_dir, "weights-*.npz"))
    if paths:
        idx = max(int(os.path.basename(p).split('-')[1].split('.')[0]) for p in paths)
        return os.path.join(checkpoint_dir, "weights-%d.npz" % idx)
    return os.path.join(checkpoint_dir, "weights.npz")


def load_weights(checkpoint_path):
    """Saver.restore (synthesizer.py:66-70) -> state dict."""
    import numpy as np
    from .. import tf_bundle
    if os.path.exists(checkpoint_path + '.index'):
        return tf_bundle.load_variables(checkpoint_path)
    with np.load(checkpoint_path) as f:
        return {k: f[k] for k in f.files}
