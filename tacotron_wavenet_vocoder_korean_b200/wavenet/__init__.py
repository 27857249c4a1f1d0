# Same exports as the reference's wavenet/__init__.py:2-3.
from .model import WaveNetModel
from .ops import mu_law_encode, mu_law_decode, optimizer_factory

__all__ = ["WaveNetModel", "mu_law_encode", "mu_law_decode", "optimizer_factory"]
