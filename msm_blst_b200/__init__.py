"""msm_blst_b200 — B200-native fixed-base MSM over BLS12-381 G1/G2.

The product is the C-ABI shared library `libmsm_b200.so` (hand-written sm_100a CUDA, see csrc/ and
include/msm_b200.h). This package is the thin host-side mirror of the reference driver's interface
(main_p1.cpp / main_p2.cpp) over that ABI, used by tests and bench.py. There is no CPU fallback: importing
works without a GPU, every compute call needs one and fails loudly otherwise.
"""
from .api import (  # noqa: F401
    LIB_PATH,
    MsmB200Error,
    MsmContext,
    affine_serialize,
    build_library,
    config_lookup,
    host_bucket_set,
    host_bucket_set_check,
    host_digit_table,
    lib,
    measure_peaks,
    test_field_op,
    test_point_op,
)

__all__ = [
    "LIB_PATH", "MsmB200Error", "MsmContext", "affine_serialize", "build_library", "config_lookup", "host_bucket_set", "host_bucket_set_check", "host_digit_table", "lib", "measure_peaks",
    "test_field_op", "test_point_op",
]
