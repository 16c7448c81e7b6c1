"""obmd-lammps-extension_b200: B200-native OBMD-DPD hot path.

Thin ctypes binding of the C ABI in include/obmd_b200.h (libobmd_b200.so, hand-written
sm_100a CUDA).  The binding is plumbing for tests and bench.py; the product is the shared
library and the C++ host layer (host/).  There is no CPU fallback: if the library is missing
or no B200 is visible, construction fails loudly.

Import as ``import obmd_b200`` (root shim) or
``importlib.import_module("obmd-lammps-extension_b200")``.
"""
from .binding import (ABI_VERSION, FixArgs, FixState, ObmdB200, ObmdError, PreExchangeReport, Trial,
                      default_regions, lib, lib_path, parse_fix_obmd, slab_bounds, slab_columns,
                      slab_mask)

__all__ = ["ABI_VERSION", "FixArgs", "FixState", "ObmdB200", "ObmdError", "PreExchangeReport",
           "Trial", "default_regions", "lib", "lib_path", "parse_fix_obmd", "slab_bounds", "slab_columns",
           "slab_mask"]
