"""The carry logic of the dedicated Montgomery squaring (csrc/fp.cuh: fp_sqr_inline), checked on the CPU through its
limb-level model: result == x^2 * R^-1 mod p and no dropped carry, on random and structured inputs."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import fp_sqr_model


def test_squaring_model_matches_montgomery_square():
    assert fp_sqr_model.check(400, seed=7) > 400
