"""clstm_b200 -- B200-native (sm_100a) replacement for the hot path of tmbdev/clstm.

This Python package is only a ctypes view of the C ABI in include/clstm_b200.h (used by tests/ and bench.py).
The product is libclstm_b200.so (clstm_b200/csrc) plus the C++ host mirror of the reference interface
(clstm_b200/host).  There is no CPU fallback: importing works anywhere, creating a net needs a B200.
"""
from ._ffi import Net, lib, build, LIB_PATH, Error, EXPORTS, pinned_array, selftest_lstm  # noqa: F401
