# coding: utf-8
"""text/symbols.py:12-17 of the reference: the Korean symbol table is the model's vocabulary."""
from .korean import ALL_SYMBOLS, PAD, EOS

symbols = ALL_SYMBOLS
