# Same exports as the reference's tacotron/__init__.py:1-13.
import os
from glob import glob

from .tacotron import Tacotron


def create_model(hparams):
    return Tacotron(hparams)


def get_most_recent_checkpoint(checkpoint_dir):
    """tacotron/__init__.py:11-13 picks the newest TF checkpoint; here the interchange file is weights-<step>.npz
    (or weights.npz) next to params.json, keyed by the TF variable names."""
    paths = glob(os.path.join(checkpoint_dir, "weights-*.npz"))
    if paths:
        idx = max(int(os.path.basename(p).split('-')[1].split('.')[0]) for p in paths)
        return os.path.join(checkpoint_dir, "weights-%d.npz" % idx)
    return os.path.join(checkpoint_dir, "weights.npz")
