"""Drop-in for the reference's `module.tinyvc` package (module/tinyvc/__init__.py:1-4).

The training-only `Discriminator` (reference discriminator.py) is out of scope of the inference
path and is not provided.
"""
from .decoder import Decoder
from .encoder import Encoder
from .feature_retrieval import match_features

__all__ = ["Encoder", "Decoder", "match_features"]
