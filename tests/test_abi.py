"""CPU tests (-m "not gpu"): the C-ABI library loads without a GPU, exports every symbol include/msm_b200.h
declares, its host-only parameter layer matches the oracle, and compute entry points fail loudly (no fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import oracle_lib as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "msm_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(msmb200_[a-zA-Z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(product_lib):
    L = product_lib.lib()
    names = declared_symbols()
    assert len(names) >= 30
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, missing


def test_config_lookup_matches_oracle(product_lib, oracle_built):
    for name in ("8", "9", "10", "11", "12", "13", "14", "15", "16", "16_beta", "17", "17_beta", "18", "19", "20",
                 "20_beta", "21"):
        c = product_lib.config_lookup(name)
        oc = O.config(name)
        for k in ("n_exp", "e", "h", "a", "d", "bsize", "e_bgmw", "h_bgmw"):
            assert getattr(c, k) == oc[k], (name, k)
        assert product_lib.lib().msmb200_pippenger_window_size(1 << c.n_exp) == oc["window"]
    with pytest.raises(KeyError):
        product_lib.config_lookup("22")


def test_host_bucket_set_and_digit_table_match_oracle(product_lib, oracle_built):
    for name in ("8", "10", "11", "13", "16_beta"):
        c = O.config(name)
        oc = O.OracleCtx(1, name, n=2)
        B = product_lib.host_bucket_set(c["e"], c["a"])
        assert len(B) == c["bsize"]
        assert (B == oc.bucket_set()).all()
        assert (product_lib.host_digit_table(c["e"], c["a"]) == oc.hash_table()).all()


def test_host_bucket_set_sizes_all_configs(product_lib, golden):
    for key, size in golden["kat_appc"]["bsize"].items():
        e, a = map(int, key.split(","))
        B = product_lib.host_bucket_set(e, a)
        assert len(B) == size
        assert int(np.diff(B).max()) == 6 and B[0] == 0 and B[1] == 1


def _py_bucket_set(q, ah):
    """Independent restatement of construct_bucket_set (reference main_bucket_set_construction.cpp:39-72) with Python sets."""
    def w(i, p):
        c = 0
        while i % p == 0:
            i //= p
            c += 1
        return c
    B = {0, 1} | {i for i in range(2, q // 2 + 1) if (w(i, 2) + w(i, 3)) % 2 == 0}
    for i in range(q // 4, q // 2):
        if i in B and (q - 2 * i) in B:
            B.discard(q - 2 * i)
    for i in range(q // 6, q // 4):
        if i in B and (q - 3 * i) in B:
            B.discard(q - 3 * i)
    B |= {i for i in range(1, ah + 2) if (w(i, 2) + w(i, 3)) % 2 == 0}
    return sorted(B)


def _py_check(B, q, ah):
    """check_bucket_set_validity (:74-113) and max_gap_in_bucket_set (:115-122)."""
    lead = {0} | {m * b for b in B for m in (1, 2, 3) if m * b <= ah + 1}
    cov = {0}
    for b in B:
        for m in (1, 2, 3):
            if m * b <= q:
                cov |= {m * b, q - m * b}
    missing = sorted(set(range(q + 1)) - cov)
    return lead == set(range(ah + 2)), not missing, (missing[0] if missing else -1), max(y - x for x, y in zip(B, B[1:]))


def test_parameter_search_check_all_reference_configs(product_lib, golden):
    """the validity / max-gap check of the reference's parameter tool on all 17 shipped (e, a) pairs."""
    for key, size in golden["kat_appc"]["bsize"].items():
        e, a = map(int, key.split(","))
        r = product_lib.host_bucket_set_check(1 << e, a)
        assert r == {"valid": True, "size": size, "max_gap": 6, "leading_ok": True, "first_uncovered": -1}, (key, r)


def test_parameter_search_check_other_radices(product_lib):
    """Radices and leading terms the reference does not ship (including non powers of two and sets that FAIL the check),
    against an independent Python restatement of construct_bucket_set + check_bucket_set_validity."""
    seen_invalid = 0
    for q in (16, 32, 64, 100, 1000, 1024, 3 * 512, 1 << 13, 10000):
        for a in (0, 3, 7, 40, 231):
            if a + 1 > q // 2:
                continue
            B = _py_bucket_set(q, a)
            lead, cover, first, gap = _py_check(B, q, a)
            r = product_lib.host_bucket_set_check(q, a)
            assert (r["size"], r["max_gap"], r["leading_ok"], r["first_uncovered"], r["valid"]) == (len(B), gap, lead, first, lead and cover), (q, a, r)
            seen_invalid += not r["valid"]
    with pytest.raises(product_lib.MsmB200Error):
        product_lib.host_bucket_set_check(15, 1)
    assert product_lib.host_bucket_set_check(1 << 13, 231)["valid"]
    print("invalid parameter pairs found:", seen_invalid)


def test_blst_register_table_needs_a_gpu(product_lib):
    """The int-returning shim entry point reports a missing CUDA device instead of aborting (the void blst-named shims abort)."""
    import ctypes as C

    import torch

    if torch.cuda.is_available():
        pytest.skip("CUDA device present")
    t = np.zeros((16, 96), dtype=np.uint8)
    L = product_lib.lib()
    assert L.msmb200_blst_register_table(1, t.ctypes.data_as(C.c_void_p), 16) != 0
    assert L.msmb200_blst_register_table(3, t.ctypes.data_as(C.c_void_p), 16) != 0     # bad group
    assert L.msmb200_blst_register_table(1, None, 16) != 0
    assert L.msmb200_blst_last_call_ms(7) < 0


def test_no_cpu_fallback(product_lib):
    """Without a CUDA device every compute entry point returns an error; with one, bad arguments do."""
    import torch

    L = product_lib.lib()
    if not torch.cuda.is_available():
        with pytest.raises(product_lib.MsmB200Error):
            product_lib.MsmContext(1, "10")
        a = np.zeros((4, 6), dtype=np.uint64)
        with pytest.raises(product_lib.MsmB200Error):
            product_lib.test_field_op(1, 0, a, a)
    h = C.c_void_p()
    cfg = product_lib.config_lookup("10")
    assert L.msmb200_ctx_create(C.byref(h), 3, C.byref(cfg), 1024, 0) == -1  # bad group
    assert L.msmb200_ctx_create(C.byref(h), 1, C.byref(cfg), 0, 0) == -1  # no points
    assert b"bad arguments" in L.msmb200_last_error(None)


def test_product_does_not_touch_the_oracle():
    """The product path must not import, link or execute anything under oracle/."""
    pkg = os.path.join(ROOT, "msm_blst_b200")
    for dirpath, _, files in os.walk(pkg):
        if os.path.basename(dirpath) == "build":
            continue
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".hpp", ".h", ".cpp")) or fn == "Makefile":
                text = open(os.path.join(dirpath, fn), errors="ignore").read()
                assert "oracle" not in text.lower().replace("test oracle", ""), os.path.join(dirpath, fn)
    text = open(os.path.join(ROOT, "include", "msm_b200.h")).read()
    assert "oracle" not in text.lower()
