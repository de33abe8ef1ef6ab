"""msmd_b200 — B200-native (sm_100a) implementation of MSMD's speech-to-face hot path.

Python here is host plumbing that keeps the reference's call surface; all compute runs in
hand-written CUDA kernels behind the C ABI in include/msmd_b200.h (libmsmd_b200.so).
There is no CPU fallback: importing works without a GPU, calling does not.
"""
from . import _lib  # noqa: F401

__all__ = ['_lib']
