#  coding: utf-8
"""Drop-in for the reference's top-level generate.py (same CLI); see
tacotron_wavenet_vocoder_korean_b200/generate.py."""
import time

from tacotron_wavenet_vocoder_korean_b200.generate import main

if __name__ == '__main__':
    s = time.time()
    main()
    print(time.time() - s, 'sec')
