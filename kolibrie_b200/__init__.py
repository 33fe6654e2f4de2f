"""kolibrie_b200 — B200-native (sm_100a) replacement for Kolibrie's data-parallel hot path.

The product is the C-ABI shared library `libkolibrie_b200.so` (include/kolibrie_b200.h, include/cudajoin.h) built from
kolibrie_b200/csrc/*.cu. This package is the thin host side used where the reference's own toolchain (Rust) is absent:
ctypes binding (capi), seeded synthetic data (datagen), the Python mirror of the reference's operator / reasoner interfaces
(engine), and the one-process-per-GPU sharding helpers (dist).

Importing the package does not load the library; the first use does, and fails loudly if it is missing — there is no CPU path.
"""
from . import capi  # noqa: F401

__all__ = ["capi"]
__version__ = "0.1.0"
