"""bundle_adjustment_solver_b200 -- B200-native (sm_100a) bundle-adjustment engine.

Drop-in for ChanghyeonKim93/bundle_adjustment_solver's analytic LM/Schur hot path:
CUDA kernels + C-ABI in csrc/ (libba_b200.so, declared in include/ba_b200.h), C++ drop-in classes
in include/ba_b200/, and this Python mirror of the same interface.  No CPU fallback.
"""
from . import scenes  # noqa: F401
from .capi import BaError, Options, PoseOnlyOptions, default_options  # noqa: F401


def build(force=False, verbose=False):
    from . import _build
    return _build.build(force=force, verbose=verbose)
