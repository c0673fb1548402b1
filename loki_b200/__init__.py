"""loki_b200 -- B200-native (sm_100a) Vlasov right-hand-side path for LLNL/LOKI.

The product is the C-ABI shared library `libloki_b200.so` (include/loki_b200.h) plus the C++ host
mirror of the reference's KineticSpecies / VPSystem / RK integrator interfaces compiled into it.
This Python package is only a ctypes loader used by the tests and bench.py; it contains no compute
and no CPU fallback: if the CUDA library is missing or no GPU is present, compute calls raise.
"""
from .capi import (  # noqa: F401
    Geom, Accel, Inflow, RkUpdate, StageMoments, LokiError, lib, load, library_path,
)
