#!/usr/bin/env python
"""Root-level entry point with the reference's name: python train_vocoder.py --data_dir ... (see the package module)."""
from tacotron_wavenet_vocoder_korean_b200.train_vocoder import main

if __name__ == '__main__':
    main()
    print('Done')
