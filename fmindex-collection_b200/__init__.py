"""fmb200 -- B200-native batched FM-index search (hot path of SGSSGene/fmindex-collection).

The product is `libfmb200.so` (hand-written sm_100a CUDA behind the C-ABI of include/fmb200.h).  This package
is the thin Python host side used by tests and bench.py: a ctypes binding (`capi`), the search-scheme
generators (`schemes`), synthetic workloads (`synth`).  There is no CPU fallback anywhere in here.
"""
from . import build as _build  # noqa: F401
from .capi import (  # noqa: F401
    FmbError,
    Index,
    Queries,
    Results,
    device_count,
    lib,
    lib_path,
)
from . import multi, schemes, synth  # noqa: F401
