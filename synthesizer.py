# coding: utf-8
"""Drop-in for the reference's top-level synthesizer.py (same class and CLI); see
tacotron_wavenet_vocoder_korean_b200/synthesizer.py."""
from tacotron_wavenet_vocoder_korean_b200.synthesizer import Synthesizer, main, attention_trim_index  # noqa: F401

if __name__ == '__main__':
    main()
