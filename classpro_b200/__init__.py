"""classpro_b200 -- B200-native implementation of ClassPro's per-read classification path.

The product is the C-ABI shared library ``libclasspro_b200.so`` (CUDA kernels for sm_100a + C host
code, see include/classpro_gpu.h) and the ``ClassPro`` command-line program built on it.  This
Python package only binds that library for tests and benchmarks; it contains no compute path of
its own and no CPU fallback: importing :mod:`classpro_b200.abi` fails if the library is missing.
"""
from .abi import (Model, Context, Batch, lib, LIB_PATH, CpgError, pack_reads, ST_FATAL)  # noqa: F401
