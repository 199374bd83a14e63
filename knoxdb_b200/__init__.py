"""knoxdb_b200 — B200 (sm_100a) implementation of KnoxDB's pack-engine scan path.

The product is the C-ABI shared library ``libknoxgpu.so`` (include/knoxgpu.h) built from
``csrc/``.  This Python package is plumbing for tests and benchmarks only: a ctypes binding of
the C ABI.  There is no CPU fallback — importing works on a CPU box (so the symbol table can
be checked), but every compute call needs a CUDA device and raises otherwise.
"""
from .lib import (  # noqa: F401
    KnoxError, Context, Program, Leaf, Stats, lib, library_path, ABI_SYMBOLS,
    INT64, INT32, INT16, INT8, UINT64, UINT32, UINT16, UINT8, FLOAT64, FLOAT32, BYTES,
    EQ, NE, GT, GE, LT, LE, IN, NIN, RANGE, OP_AND, OP_OR,
)
