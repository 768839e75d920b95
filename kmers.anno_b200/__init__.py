"""kmers.anno_b200 — B200 (sm_100a) engine for the k-mer annotation hot path of
SEEDtk/kmers.anno (`apply`: ApplyKmerProcessor.java:99-148).

The product is the C-ABI library `libkmeranno.so` (include/kmeranno.h) plus the C++ host
mirror of the reference's processor/reporters (`host/`, binary `bin/kmers-anno`).  This
Python layer is a thin ctypes binding used by tests and bench.py; it contains no compute
and no fallback: if the CUDA library is missing or no GPU is present it raises.
"""
from .engine import (  # noqa: F401
    Engine,
    KmerAnnoError,
    LIB_PATH,
    load_library,
    FLAG_NONE,
    FLAG_CALLED,
    FLAG_AMBIGUOUS,
    FLAG_BELOW_MIN,
)
from . import synth  # noqa: F401
