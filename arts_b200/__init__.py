"""arts_b200 — B200-native clear-sky spectral hot path for ARTS (lbl Voigt sum + rtepack Stokes chain).

The compute lives in ``arts_b200/csrc`` (hand-written CUDA for sm_100a behind the C ABI
of ``include/arts_b200.h``).  This package is the thin host mirror of the reference's
workspace methods for that path (``arts_b200.wsm``) plus synthetic inputs
(``arts_b200.synth``).  There is no CPU fallback: every compute call raises if the
CUDA library or a CUDA device is missing.
"""
from . import _abi  # noqa: F401
from ._abi import AtmPath, HostCatalog  # noqa: F401

__all__ = ["AtmPath", "HostCatalog", "_abi"]
