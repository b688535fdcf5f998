"""qpalette -- B200-native quantized-linear decode path of Q-Palette.

Host-side mirror of the reference's `lib/linear` interface over libqpalette.so (C ABI, include/qpalette.h)."""
from . import _cabi  # noqa: F401

__all__ = ["_cabi"]
