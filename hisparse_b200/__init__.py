"""hisparse_b200 -- B200-native SpMV engine behind HiSparse's host/kernel boundary.

The product is the sm_100a shared library `libhisparse_b200.so` (sources in csrc/, C ABI in
include/hisparse_b200.h) plus the C++ host mirror in host/. This Python package only binds the
C ABI for the test-suite and bench.py.
"""
from . import capi  # noqa: F401
