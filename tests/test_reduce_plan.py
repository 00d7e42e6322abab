"""Host logic of the digit-splitting bucket reduction (the static list plans walked by list_sum_kernel /
list_sum_coop_kernel, which replace POINTonE1_integrate_buckets_accumulation_d_CHES, src/multi_scalar.c:301-321, and
POINTonE1_integrate_buckets, :281-297): the plan is evaluated on 64-bit integers instead of points — no GPU — and must
reproduce sum_l value(l) * x[l] for the bucket set of every reference configuration and for dense windows."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import msm_blst_b200 as M

ALL_CONFIGS = ["8", "9", "10", "11", "12", "13", "14", "15", "16", "16_beta", "17", "17_beta", "18", "19", "20", "20_beta", "21"]


def plan_eval(values, nbw, x, res1=2, resq=3, gpw=8):
    L = M.lib()
    fn = L.msmb200_host_reduce_plan_eval
    fn.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    out = C.c_uint64(0)
    info = (C.c_uint32 * 8)()
    vp = values.ctypes.data if values is not None else None
    rc = fn(vp, nbw, res1, resq, gpw, x.ctypes.data, C.byref(out), info)
    assert rc == 0
    return out.value, list(info)


def bucket_set(cfgname):
    # the C ABI directly: msmb200_config_lookup + msmb200_host_bucket_set (both host-only)
    L = M.lib()
    class Cfg(C.Structure):
        _fields_ = [(n, C.c_int) for n in ("n_exp", "e", "h", "a", "d", "bsize", "e_bgmw", "h_bgmw")]
    cc = Cfg()
    L.msmb200_config_lookup.argtypes = [C.c_char_p, C.c_void_p]
    assert L.msmb200_config_lookup(cfgname.encode(), C.byref(cc)) == 0
    L.msmb200_host_bucket_set.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_long]
    L.msmb200_host_bucket_set.restype = C.c_long
    size = L.msmb200_host_bucket_set(cc.e, cc.a, None, 0)
    out = np.zeros(size, dtype=np.int32)
    assert L.msmb200_host_bucket_set(cc.e, cc.a, out.ctypes.data, size) == size
    return out, cc


@pytest.mark.parametrize("cfgname", ALL_CONFIGS)
def test_plan_reproduces_weighted_sum_for_every_reference_bucket_set(cfgname):
    B, cc = bucket_set(cfgname)
    if cc.bsize:
        assert len(B) == cc.bsize
    rng = np.random.default_rng(len(B))
    x = rng.integers(0, 2**64, size=len(B), dtype=np.uint64)
    x[rng.integers(1, len(B), size=len(B) // 7)] = 0  # empty buckets
    exp = int(sum(int(b) * int(v) for b, v in zip(B[1:], x[1:])) % 2**64)
    for res1, resq, gpw in ((2, 3, 8), (3, 2, 4)):  # G1-like and G2-like residency / lane-group geometry
        got, info = plan_eval(B, len(B), x, res1, resq, gpw)
        assert got == exp, (cfgname, res1)
        c_lo, nbits, slice1, nslices, nlists = info[0], info[1], info[2], info[3], info[4]
        assert nbits == int(B[-1]).bit_length() and 1 <= c_lo <= nbits
        assert nslices <= 32 * res1 * 4 * 148 or slice1 == 2  # stage 1 is one wave of resident lanes
        assert info[5] in (1, 2, 4, 8) and info[5] <= gpw and info[7] <= gpw


@pytest.mark.parametrize("nbw", [2, 3, 9, 129, 4097, (1 << 17) + 1, (1 << 21) + 1])
def test_plan_dense_windows(nbw):
    """value == local index: BGMW95 (2^(e'-1) + 1 buckets), blst's Pippenger windows, blst_pNs_add / mult_wbits (nbw = 2)."""
    rng = np.random.default_rng(nbw)
    x = rng.integers(0, 2**64, size=nbw, dtype=np.uint64)
    exp = int(sum(i * int(v) for i, v in enumerate(x.tolist())) % 2**64)
    got, info = plan_eval(None, nbw, x)
    assert got == exp
    assert info[1] == (nbw - 1).bit_length()
