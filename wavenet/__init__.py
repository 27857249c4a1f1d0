# Top-level `wavenet` package with the reference's exports (wavenet/__init__.py:2-3), so
# `from wavenet import WaveNetModel, mu_law_decode, mu_law_encode` (generate.py:31) keeps working.
from tacotron_wavenet_vocoder_korean_b200.wavenet import WaveNetModel, mu_law_encode, mu_law_decode, optimizer_factory  # noqa: F401
