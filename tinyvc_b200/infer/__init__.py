"""Drop-in for the reference's `module.infer` package (module/infer/__init__.py:1-2)."""
from .generator import Generator
from .stream import StreamInfer, BatchedStreamInfer

__all__ = ["Generator", "StreamInfer", "BatchedStreamInfer"]
