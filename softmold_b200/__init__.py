"""softmold_b200 -- B200-native replacement of the SoftMold `MD` timestep.

The product is `libsoftmold_b200.so` (C ABI, include/softmold_b200.h: hand-written sm_100a CUDA + the `.mpd` file
boundary) and the `MD_b200` host driver.  This package only binds the C ABI for tests and bench.py.
There is no CPU fallback: `capi.lib()` raises when the library has not been built, and every compute call raises
when no B200-class GPU is usable."""
from .capi import (Context, Mpd, SoftMoldError, lib, LIB_PATH, SYMBOLS, MASK_ALL, MASK_ALL_MOLECULES, MASK_LANGEVIN,  # noqa: F401
                   TERM_PAIR, TERM_CHAIN, TERM_BOND, TERM_BEND, TERM_BEAD, TERM_BALL, TERM_FIELD, TERM_NANOCORE, NTERMS, NOISE_PHILOX,
                   NOISE_EXTERNAL, MOL_BOND, MOL_BEND, MOL_CHAIN, MOL_BEAD, MOL_BALL, MOL_BOUNDARY, MOL_FLOATING_BASE, MOL_ZTORQUE,
                   MOL_ZPOWERPOTENTIAL, MOL_NANOCORE, MOL_IGNORED, PHASES)
