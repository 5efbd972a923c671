"""supermc_b200 -- B200-native per-event initial-condition hot path of superMC.

The product is the C-ABI shared library ``libsupermc_b200.so`` (hand-written sm_100a CUDA behind
``include/supermc_b200.h``) plus the C++ host driver that mirrors the reference's MakeDensity modes.
This package only holds the build recipe and a thin ctypes binding used by tests, bench.py and the
multi-GPU launcher.  There is no CPU fallback: importing works anywhere, computing needs a B200.
"""
from .capi import (Params, Constants, EventOut, EventIn, Context, SmcError, lib, build_library,
                   RUN_MOMENTS, RUN_KEEP_RHO, RUN_THICKNESS, RUN_RHO_BINARY, RUN_SPECTATORS, RUN_LISTS,
                   GRID_RHO, GRID_TA1, GRID_TA2, GRID_RHO_BINARY, GRID_SPEC_A, GRID_SPEC_B)
