"""The branch-free safegcd inversion of the batch-affine kernels (csrc/inv.cuh), compiled for the HOST from the same
source the CUDA kernels use, against Python's modular inverse: out == x^-1 * 2^768 mod p (Montgomery-form inverse of a
Montgomery-form input), 0 -> 0, edge values, and the round count stays below the guard."""
import ctypes
import os
import random
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
P = 0x1a0111ea397fe69a4b1ba7b6434bacd764774b84f38512bf6730d2a0f6b0f6241eabfffeb153ffffb9feffffffffaaab
R = 1 << 384


def _lib():
    src = os.path.join(HERE, "native", "inv_model.cpp")
    out = os.path.join(HERE, "native", "libinv_model.so")
    hdr = os.path.join(HERE, "..", "msm_blst_b200", "csrc", "inv.cuh")
    if not os.path.exists(out) or os.path.getmtime(out) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.check_call(["/usr/bin/g++", "-O2", "-std=c++17", "-x", "c++", "-shared", "-fPIC", "-o", out, src])
    lib = ctypes.CDLL(out)
    lib.inv_model_batch.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
    lib.inv_model_batch.restype = ctypes.c_int
    return lib


def _words(x):
    return [(x >> (32 * i)) & 0xffffffff for i in range(12)]


def test_safegcd_inverse_matches_python():
    rng = random.Random(11)
    xs = [0, 1, 2, 3, P - 1, P - 2, (P + 1) // 2, (P - 1) // 2, R % P, (R * R) % P, 1 << 380, (1 << 381) - 1 - ((1 << 381) - 1 >= P) * 0]
    xs = [x % P for x in xs]
    xs += [(1 << k) % P for k in range(0, 381, 7)] + [P - (1 << k) for k in range(0, 380, 11)]
    xs += [rng.randrange(P) for _ in range(20000)]
    xs += [rng.randrange(1 << rng.randrange(1, 381)) for _ in range(2000)]  # short values
    a = np.array([_words(x) for x in xs], dtype=np.uint32)
    out = np.zeros_like(a)
    rounds = _lib().inv_model_batch(a.ctypes.data, out.ctypes.data, len(xs))
    assert 20 <= rounds <= 30, rounds  # the guard (31) is never the limiter
    for x, o in zip(xs, out):
        got = sum(int(w) << (32 * i) for i, w in enumerate(o))
        exp = 0 if x == 0 else pow(x, -1, P) * R * R % P
        assert got == exp, hex(x)
