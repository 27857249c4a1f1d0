# Top-level `text` package with the reference's entry points (text/__init__.py:38,76).
from tacotron_wavenet_vocoder_korean_b200.text import text_to_sequence, sequence_to_text, prepare_inputs  # noqa: F401
