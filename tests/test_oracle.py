"""CPU tests (-m "not gpu"): pin the ORACLE (oracle/msm_oracle.cpp) against the reference's golden vectors
(tests/golden/golden.json, generated from the compiled reference) and, when oracle/_ref is present, against the
compiled reference itself."""
import ctypes as C
import hashlib

import numpy as np
import pytest

import oracle_lib as O


def h2a(hexstr, dtype=np.uint64, shape=None):
    a = np.frombuffer(bytes.fromhex(hexstr), dtype=dtype).copy()
    return a.reshape(shape) if shape else a


def test_field_golden(golden, oracle_built):
    o = O.oracle()
    f = golden["field"]
    a = h2a(f["a"], shape=(-1, 6))
    b = h2a(f["b"], shape=(-1, 6))
    n = a.shape[0]
    for name, op in (("mul", 0), ("sqr", 1), ("add", 2), ("sub", 3), ("mul_by_3", 5), ("inverse", 6)):
        out = np.empty_like(a)
        assert o.oracle_fp_op(op, O.ptr(a), O.ptr(b), O.ptr(out), n) == 0
        assert out.tobytes().hex() == f[name], name
    a2, b2 = a.reshape(-1, 12), b.reshape(-1, 12)
    for name, op in (("fp2_mul", 0), ("fp2_sqr", 1), ("fp2_add", 2), ("fp2_sub", 3)):
        out = np.empty_like(a2)
        assert o.oracle_fp2_op(op, O.ptr(a2), O.ptr(b2), O.ptr(out), a2.shape[0]) == 0
        assert out.tobytes().hex() == f[name], name


@pytest.mark.skipif(not O.has_ref(), reason="compiled reference (oracle/_ref) not present")
def test_field_vs_compiled_reference(oracle_built):
    o, b = O.oracle(), O.blst_ref()
    rng = np.random.default_rng(5)
    n = 500
    a = rng.integers(0, 2**64, size=(n, 6), dtype=np.uint64)
    bb = rng.integers(0, 2**64, size=(n, 6), dtype=np.uint64)
    a[:, 5] &= np.uint64(0x0FFFFFFFFFFFFFFF)
    bb[:, 5] &= np.uint64(0x0FFFFFFFFFFFFFFF)
    for op, fn in ((0, "blst_fp_mul"), (2, "blst_fp_add"), (3, "blst_fp_sub")):
        out, ref = np.empty_like(a), np.empty_like(a)
        o.oracle_fp_op(op, O.ptr(a), O.ptr(bb), O.ptr(out), n)
        f = getattr(b, fn)
        for i in range(n):
            f(C.c_void_p(ref[i].ctypes.data), C.c_void_p(a[i].ctypes.data), C.c_void_p(bb[i].ctypes.data))
        assert (out == ref).all(), fn


@pytest.mark.skipif(not O.has_ref(), reason="compiled reference (oracle/_ref) not present")
def test_point_ops_vs_compiled_reference(oracle_built):
    """xyzz_dadd_affine / xyzz_dadd / add_or_double incl. doubling, cancellation and infinity branches."""
    o, b = O.oracle(), O.blst_ref()
    for g in (1, 2):
        ab, jb, xb = O.AFF_BYTES[g], O.JAC_BYTES[g], O.XYZZ_BYTES[g]
        oc = O.OracleCtx(g, "10", n=16)
        oc.init_fix_points()
        pts = oc.points().reshape(16, ab)
        acc = np.zeros(xb, dtype=np.uint8)
        seq = [(0, 0), (1, 0), (1, 1), (0, 1), (2, 0), (2, 0), (3, 1), (3, 0), (4, 0)]  # P, +Q, -Q, -P (->inf), dbl ...
        for idx, sign in seq:
            exp = np.zeros(xb, dtype=np.uint8)
            getattr(b, "blst_p%dxyzz_dadd_affine" % g)(O.ptr(exp), O.ptr(acc), O.ptr(pts[idx].copy()), C.c_uint32(sign))
            got = np.zeros(xb, dtype=np.uint8)
            fl = np.array([sign], dtype=np.uint8)
            o.oracle_point_op(g, 2, O.ptr(acc), O.ptr(pts[idx].copy()), O.ptr(fl), O.ptr(got), 1)
            # infinity: the reference leaves X,Y stale; compare ZZZ,ZZ always and X,Y when finite
            half = xb // 2
            assert (got[half:] == exp[half:]).all()
            if got[half:].any():
                assert (got == exp).all()
            acc = exp
        # xyzz + xyzz incl. doubling
        x1 = np.zeros(xb, dtype=np.uint8)
        getattr(b, "blst_p%dxyzz_dadd_affine" % g)(O.ptr(x1), O.ptr(np.zeros(xb, dtype=np.uint8)), O.ptr(pts[5].copy()), C.c_uint32(0))
        x2 = np.zeros(xb, dtype=np.uint8)
        getattr(b, "blst_p%dxyzz_dadd_affine" % g)(O.ptr(x2), O.ptr(x1), O.ptr(pts[6].copy()), C.c_uint32(0))
        for lhs, rhs in ((x1, x2), (x2, x2), (x1, x1)):
            exp, got = np.zeros(xb, dtype=np.uint8), np.zeros(xb, dtype=np.uint8)
            getattr(b, "blst_p%dxyzz_dadd" % g)(O.ptr(exp), O.ptr(lhs), O.ptr(rhs))
            o.oracle_point_op(g, 3, O.ptr(lhs), O.ptr(rhs), None, O.ptr(got), 1)
            assert (got == exp).all()
        # jacobian add-or-double and to_affine
        j1, j2 = np.zeros(jb, dtype=np.uint8), np.zeros(jb, dtype=np.uint8)
        getattr(b, "blst_p%d_from_affine" % g)(O.ptr(j1), O.ptr(pts[1].copy()))
        getattr(b, "blst_p%d_from_affine" % g)(O.ptr(j2), O.ptr(pts[2].copy()))
        for lhs, rhs in ((j1, j2), (j1, j1)):
            exp, got = np.zeros(jb, dtype=np.uint8), np.zeros(jb, dtype=np.uint8)
            getattr(b, "blst_p%d_add_or_double" % g)(O.ptr(exp), O.ptr(lhs), O.ptr(rhs))
            o.oracle_point_op(g, 0, O.ptr(lhs), O.ptr(rhs), None, O.ptr(got), 1)
            assert (got == exp).all()
            ea, ga = np.zeros(ab, dtype=np.uint8), np.zeros(ab, dtype=np.uint8)
            getattr(b, "blst_p%d_to_affine" % g)(O.ptr(ea), O.ptr(exp))
            o.oracle_point_op(g, 5, O.ptr(got), None, None, O.ptr(ga), 1)
            assert (ga == ea).all()


def test_bucket_set_kats(golden, oracle_built):
    """|B| per (e, a) = the B_SIZE constants of ches_config_files/*.h; max gap 6; full digit coverage."""
    for key, size in golden["kat_appc"]["bsize"].items():
        e, a = map(int, key.split(","))
        if e > 20:
            continue  # q = 2^22 takes ~10 s with std::set; covered by the product's host test below
        assert O.oracle().oracle_bucket_set(e, a, None, 0) == size
    for e, a in ((12, 7), (13, 231), (16, 29677)):
        assert O.oracle().oracle_bucket_set_check(e, a) == 6


def test_config_table(oracle_built):
    assert O.config("10") == dict(n_exp=10, e=13, h=20, a=231, d=6, bsize=1725, e_bgmw=12, h_bgmw=22, window=8)
    assert O.config("21")["window"] == 18 and O.config("18")["window"] == 15 and O.config("16")["window"] == 13
    for name in ("8", "9", "11", "12", "13", "14", "15", "16_beta", "17", "17_beta", "19", "20", "20_beta"):
        c = O.config(name)
        assert c["e"] * c["h"] >= 255 and c["e_bgmw"] * c["h_bgmw"] >= 255


@pytest.mark.parametrize("group", [1, 2])
def test_c10_against_reference_driver_golden(golden, oracle_built, group):
    """Oracle = reference driver (main_p{1,2}.cpp, config 10) on fixed points, tables, digits and all 4 methods."""
    gd = golden["c10"][str(group)]
    oc = O.OracleCtx(group, "10", threads=4)
    oc.init_fix_points()
    assert hashlib.sha256(oc.points().tobytes()).hexdigest() == gd["sha256_fix_points"]
    assert hashlib.sha256(oc.bucket_set().tobytes()).hexdigest() == gd["sha256_bucket_set"]
    oc.build_table(0)
    oc.build_table(1)
    assert hashlib.sha256(oc.table(0).tobytes()).hexdigest() == gd["sha256_table_3nh"]
    assert hashlib.sha256(oc.table(1).tobytes()).hexdigest() == gd["sha256_table_bgmw95"]
    sc = O.gen_scalars(1, oc.n)
    for i, dg in enumerate(gd["digits_seed1"]):
        m, b = oc.digits(0, sc[i])
        assert m.tolist() == dg["m"] and b.tolist() == dg["b"]
        qh, _ = oc.digits(1, sc[i])
        assert qh.tolist() == dg["qhalf"]
    seeds = (1, 2, 3) if group == 1 else (1,)
    for seed in seeds:
        sc = O.gen_scalars(seed, oc.n)
        assert not oc.hits_bug(sc)
        for method in (1, 2, 3, 4):
            r = oc.msm(method, sc, faithful_bug=(method <= 2))
            assert O.serialize(group, r).hex() == gd["msm"][str(seed)], (seed, method)
        cf, _ = O.closed_form(group, sc)
        assert O.serialize(group, cf).hex() == gd["msm"][str(seed)]


def test_appc_known_answers_closed_form(golden, oracle_built):
    """SURVEY App. C at every BASELINE size through the closed form (sum s_i 2^(i+1) mod r) * G."""
    k = golden["kat_appc"]
    sc = O.gen_scalars(1, 1 << 21)
    assert "%064x" % O.scalars_to_ints(sc[:1])[0] == k["scalar0_seed1"]
    for g, nexp, key in ((1, 10, "g1_n10"), (1, 16, "g1_n16"), (1, 21, "g1_n21"), (2, 18, "g2_n18")):
        cf, kk = O.closed_form(g, sc[: 1 << nexp])
        assert "%064x" % kk == k[key + "_k"] if key + "_k" in k else True
        assert O.serialize(g, cf).hex() == k[key]
    oc = O.OracleCtx(1, "10", n=4)
    oc.init_fix_points()
    assert O.serialize(1, oc.points()[:96]).hex() == k["g1_p0"]


@pytest.mark.parametrize("group", [1, 2])
def test_pippenger_unstructured_golden(golden, oracle_built, group):
    """blst_pNs_mult_pippenger of the compiled reference on non-structured points (incl. infinity, zero scalar)."""
    gd = golden["pippenger_unstructured"][str(group)]
    n = gd["n"]
    pts = np.frombuffer(bytes.fromhex(gd["points"]), dtype=np.uint8).copy()
    sc = np.frombuffer(bytes.fromhex(gd["scalars"]), dtype=np.uint64).reshape(n, 4).copy()
    oc = O.OracleCtx(group, "8", n=n)
    oc.set_points(pts)
    r = oc.msm(4, sc)
    assert O.serialize(group, r).hex() == gd["result"]
    naive = np.zeros(O.AFF_BYTES[group], dtype=np.uint8)
    O.oracle().oracle_naive_msm(group, O.ptr(pts), O.ptr(sc), 40, O.ptr(naive))
    oc2 = O.OracleCtx(group, "8", n=40)
    oc2.set_points(pts[: 40 * O.AFF_BYTES[group]])
    assert (oc2.msm(4, sc[:40]) == naive).all()


def test_last_element_bug_policy(oracle_built):
    """the reference drops the last point when the penultimate digit maps to bucket 0.
    The oracle can reproduce it (faithful_bug) and detect inputs that hit it; the product computes the true sum."""
    oc = O.OracleCtx(1, "10", n=8)
    oc.init_fix_points()
    oc.build_table(0)
    sc = O.gen_scalars(3, 8)
    # force digit h-2 of the last scalar to 0 and digit h-1 to non-zero: bits [13*18, 13*19) cleared
    v = O.scalars_to_ints(sc[7:8])[0]
    v &= ~(((1 << 26) - 1) << (13 * 17))  # digits 17 and 18 zero: no carry can reach digit 18
    v |= 1 << (13 * 19)
    v &= (1 << 255) - 1
    sc[7] = [(v >> (64 * i)) & (2**64 - 1) for i in range(4)]
    assert oc.hits_bug(sc)
    true_sum, _ = O.closed_form(1, sc)
    assert (oc.msm(1, sc, faithful_bug=False) == true_sum).all()
    assert not (oc.msm(1, sc, faithful_bug=True) == true_sum).all()
    assert (oc.msm(3 if False else 4, sc) == true_sum).all()


@pytest.mark.skipif(not O.has_ref(), reason="compiled reference (oracle/_ref) not present")
def test_last_element_bug_is_real_in_the_reference(oracle_built):
    """The compiled reference driver (config 10) really drops the last point on such inputs (methods 1-2),
    while its methods 3-4 return the true sum; the oracle's faithful mode reproduces methods 1-2 byte for byte."""
    d = O.refdrv(1)
    n = 1024
    sc = O.gen_scalars(4, n)
    v = O.scalars_to_ints(sc[n - 1:n])[0]
    v &= ~(((1 << 26) - 1) << (13 * 17))  # digits 17 and 18 zero: no carry can reach digit 18
    v |= 1 << (13 * 19)
    v &= (1 << 255) - 1
    sc[n - 1] = [(v >> (64 * i)) & (2**64 - 1) for i in range(4)]
    outs = {}
    for m in (1, 2, 3, 4):
        ob = np.zeros(96, dtype=np.uint8)
        d.refdrv_msm(m, O.ptr(sc), None, O.ptr(ob))
        outs[m] = ob.tobytes().hex()
    true_sum, _ = O.closed_form(1, sc)
    assert outs[3] == outs[4] == O.serialize(1, true_sum).hex()
    assert outs[1] == outs[2] != outs[3]
    oc = O.OracleCtx(1, "10", threads=4)
    oc.init_fix_points()
    oc.build_table(0)
    assert oc.hits_bug(sc)
    assert O.serialize(1, oc.msm(1, sc, faithful_bug=True)).hex() == outs[1]
    assert O.serialize(1, oc.msm(2, sc, faithful_bug=True)).hex() == outs[2]
