"""aboria_b200 — B200-native ordered cell list + sparse kernel operator product,
a drop-in for Aboria's CellListOrdered / create_sparse_operator hot path.

The product is libabr.so (hand-written sm_100a CUDA behind the C-ABI of
include/abr.h).  This package is the Python mirror of the reference interface
used by the tests and bench; the C++ mirror is include/aboria_b200/Aboria.h.
"""
from . import kernels  # noqa: F401
from ._lib import AbrError, LIB_PATH  # noqa: F401


def __getattr__(name):
    if name in ("Particles", "SparseOperator", "create_sparse_operator", "Query", "BlockOperator", "ZeroOperator",
                "create_block_operator", "create_zero_operator", "accumulate_within_distance"):
        from . import particles

        return getattr(particles, name)
    raise AttributeError(name)
