"""chinium_b200 -- B200-native direct-SCF J/K (Fock) build engine, drop-in for Chinium's Int4C2E path.

The product is the C-ABI shared library `libchinium_fock.so` (include/chinium_fock.h) built from
chinium_b200/csrc (CUDA, sm_100a).  This package is the thin Python mirror of the reference's
`Int4C2E` interface used by the tests and the benchmark; it loads the library and FAILS LOUDLY if
it is missing -- there is no CPU fallback anywhere in the product path.
"""
from .fock import Int4C2E, Int2C1E, FockEngineError, load_library, LIB_PATH  # noqa: F401
