"""B200-native OFDFT hot path with PROFESS-AD's Python API.

    from profess_ad_b200.system import System            # or:  from professad.system import System
    from profess_ad_b200.functionals import WangGovindCarter99, Hartree, ...

The compute path is hand-written sm_100a CUDA behind a C ABI (include/professad_b200.h); torch is
the device-memory / stream / autograd plumbing.  No CPU fallback.
"""
__version__ = '0.1.0'

from . import _native  # noqa: F401


def build(force=False, verbose=False):
    from .build import build as _b
    return _b(force=force, verbose=verbose)
