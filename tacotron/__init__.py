# Top-level `tacotron` package with the reference's exports (tacotron/__init__.py), so
# `from tacotron import create_model, get_most_recent_checkpoint` (synthesizer.py:19) keeps working.
from tacotron_wavenet_vocoder_korean_b200.tacotron import Tacotron, create_model, get_most_recent_checkpoint  # noqa: F401
