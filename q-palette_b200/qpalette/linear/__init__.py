"""lib/linear-compatible module API (reference: lib/linear/__init__.py exports)."""
from .comb_linear import CombLinearTCQ, CombtLinearTCQ
from .incoherent_linear import (IncoherentLinear, IncoherentMLP, IncoherentSdpaAttention, StaticKVCache, make_linear)
from .tcq_linear import QTIPLinearTCQ
from .vq_linear import VQLinearPackSIMT, VQLinearPackTensorCore

__all__ = ["QTIPLinearTCQ", "CombLinearTCQ", "CombtLinearTCQ", "VQLinearPackTensorCore", "VQLinearPackSIMT",
           "IncoherentLinear", "IncoherentMLP", "IncoherentSdpaAttention", "StaticKVCache", "make_linear"]
