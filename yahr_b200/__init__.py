"""yahr_b200 -- B200-native implementation of yahr's per-pixel render loop.

The product is `libyahr_b200.so` (hand-written sm_100a CUDA kernels behind the C ABI declared in
include/yahr_b200.h).  This package is the thin Python plumbing around it: a ctypes binding
(`yahr_b200.api`), the synthetic scene generators of the BASELINE configs (`yahr_b200.scenes`),
the reference's tile arithmetic (`yahr_b200.tiles`) and the one-process-per-GPU driver
(`yahr_b200.dist`).  There is no CPU fallback: every render call goes through the CUDA library
and raises if it is missing.
"""
__version__ = "0.1.0"
