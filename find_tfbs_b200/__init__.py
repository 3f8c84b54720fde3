"""B200-native find-tfbs hot path: haplotype build -> PWM scan -> per-haplotype TFBS counts.

The product is the C-ABI library libtfbs_b200.so (include/tfbs.h, hand-written sm_100a kernels in csrc/) and the
C++ driver csrc/driver (the reference's CLI).  This Python package only binds the library for tests and bench.py.
"""
from .binding import Block, Context, PatternSet, TfbsError, build, lib  # noqa: F401
