"""GPU parity tests (-m gpu): the CUDA path, called through the C ABI (libmsm_b200.so), against the oracle, the
golden vectors generated from the compiled reference, and SURVEY App. C known answers. Bit-exact everywhere."""
import ctypes as C
import hashlib

import numpy as np
import pytest

import oracle_lib as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def M(product_lib):
    import torch

    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return product_lib


def rand_fp(rng, n):
    a = rng.integers(0, 2**64, size=(n, 6), dtype=np.uint64)
    a[:, 5] &= np.uint64(0x0FFFFFFFFFFFFFFF)
    return a


def edge_fp():
    vals = [0, 1, 2, O.P_MOD - 1, O.P_MOD - 2, (O.P_MOD - 1) // 2, 2**380, 2**381 - 1 - (2**381 - 1 - O.P_MOD + 1) if False else O.P_MOD // 3]
    return np.array([[(v >> (64 * i)) & (2**64 - 1) for i in range(6)] for v in vals], dtype=np.uint64)


# ---------------------------------------------------------------- field layer (a4, a5, a6, a7)
def test_fp_ops_vs_oracle(M):
    rng = np.random.default_rng(1)
    a = np.concatenate([rand_fp(rng, 1 << 16), edge_fp(), edge_fp()[::-1]])
    b = np.concatenate([rand_fp(rng, 1 << 16), edge_fp()[::-1], edge_fp()])
    n = a.shape[0]
    for op in range(6):
        exp = np.empty_like(a)
        O.oracle().oracle_fp_op(op, O.ptr(a), O.ptr(b), O.ptr(exp), n)
        got = M.test_field_op(1, op, a, b)
        assert (got == exp).all(), "fp op %d" % op
    exp = np.empty_like(a[:300])
    x = np.ascontiguousarray(a[-300:])
    O.oracle().oracle_fp_op(6, O.ptr(x), None, O.ptr(exp), 300)
    assert (M.test_field_op(1, 6, x) == exp).all()


def test_fp_golden_from_reference(M, golden):
    f = golden["field"]
    a = np.frombuffer(bytes.fromhex(f["a"]), dtype=np.uint64).reshape(-1, 6).copy()
    b = np.frombuffer(bytes.fromhex(f["b"]), dtype=np.uint64).reshape(-1, 6).copy()
    for name, op in (("mul", 0), ("sqr", 1), ("add", 2), ("sub", 3), ("mul_by_3", 5), ("inverse", 6)):
        assert M.test_field_op(1, op, a, b).tobytes().hex() == f[name], name
    a2, b2 = a.reshape(-1, 12), b.reshape(-1, 12)
    for name, op in (("fp2_mul", 0), ("fp2_sqr", 1), ("fp2_add", 2), ("fp2_sub", 3)):
        assert M.test_field_op(2, op, a2, b2).tobytes().hex() == f[name], name


def test_fp2_ops_vs_oracle(M):
    rng = np.random.default_rng(2)
    a = rand_fp(rng, 1 << 15).reshape(-1, 12)
    b = rand_fp(rng, 1 << 15).reshape(-1, 12)
    for op in range(6):
        exp = np.empty_like(a)
        O.oracle().oracle_fp2_op(op, O.ptr(a), O.ptr(b), O.ptr(exp), a.shape[0])
        assert (M.test_field_op(2, op, a, b) == exp).all(), "fp2 op %d" % op
    x = np.ascontiguousarray(a[:200])
    exp = np.empty_like(x)
    O.oracle().oracle_fp2_op(6, O.ptr(x), None, O.ptr(exp), 200)
    assert (M.test_field_op(2, 6, x) == exp).all()


# ---------------------------------------------------------------- curve layer (a8, a9, a10, a22, a23)
@pytest.mark.parametrize("group", [1, 2])
def test_point_ops_vs_oracle_with_special_cases(M, group):
    ab, jb, xb = O.AFF_BYTES[group], O.JAC_BYTES[group], O.XYZZ_BYTES[group]
    oc = O.OracleCtx(group, "10", n=64)
    oc.init_fix_points()
    pts = oc.points().reshape(64, ab)
    o = O.oracle()
    # build XYZZ accumulators: inf, P_i, P_i + P_j ...
    n = 256
    rng = np.random.default_rng(3)
    acc = np.zeros((n, xb), dtype=np.uint8)
    addend = pts[rng.integers(0, 64, size=n)].copy()
    flags = rng.integers(0, 2, size=n).astype(np.uint8)
    for rounds in range(4):
        exp = np.zeros_like(acc)
        o.oracle_point_op(group, 2, O.ptr(acc), O.ptr(addend), O.ptr(flags), O.ptr(exp), n)
        got = M.test_point_op(group, 2, acc, addend, flags).reshape(n, xb)
        half = xb // 2
        assert (got[:, half:] == exp[:, half:]).all()
        finite = exp[:, half:].any(axis=1)
        assert (got[finite] == exp[finite]).all()
        acc = exp
        # next round: include same point (doubling), opposite (cancellation), infinity addend
        addend = pts[rng.integers(0, 64, size=n)].copy()
        flags = rng.integers(0, 2, size=n).astype(np.uint8)
        if rounds == 0:
            prev = addend.copy()
        if rounds == 1:
            addend[:32] = 0  # affine infinity
    # force P + P and P - P through xyzz_add_affine
    single = np.zeros((4, xb), dtype=np.uint8)
    first = np.stack([pts[3], pts[3], pts[4], pts[4]])
    z = np.zeros(4, dtype=np.uint8)
    s1 = np.zeros_like(single)
    o.oracle_point_op(group, 2, O.ptr(single), O.ptr(first), O.ptr(z), O.ptr(s1), 4)
    fl = np.array([0, 1, 0, 1], dtype=np.uint8)
    exp = np.zeros_like(single)
    o.oracle_point_op(group, 2, O.ptr(s1), O.ptr(first), O.ptr(fl), O.ptr(exp), 4)
    got = M.test_point_op(group, 2, s1, first, fl).reshape(4, xb)
    half = xb // 2
    assert (got[:, half:] == exp[:, half:]).all() and (got[0] == exp[0]).all() and (got[2] == exp[2]).all()
    assert not exp[1, half:].any() and not exp[3, half:].any()
    # xyzz + xyzz (general, doubling, cancellation via negated copy is covered by accumulate tests)
    lhs = acc
    rhs = np.roll(acc, 1, axis=0).copy()
    rhs[:16] = lhs[:16]  # doubling
    exp = np.zeros_like(lhs)
    o.oracle_point_op(group, 3, O.ptr(lhs), O.ptr(rhs), None, O.ptr(exp), n)
    got = M.test_point_op(group, 3, lhs, rhs).reshape(n, xb)
    finite = exp[:, half:].any(axis=1)
    assert (got[:, half:] == exp[:, half:]).all() and (got[finite] == exp[finite]).all()
    # quad-cooperative addition / doubling (coop.cuh): same formulas spread over 4 lanes, so the same bytes as the
    # one-thread versions, including infinity operands on either side, doubling and cancellation
    rhs2 = rhs.copy()
    rhs2[16:24] = 0                      # infinity addend
    lhs2 = lhs.copy()
    lhs2[24:32] = 0                      # infinity accumulator
    exp2 = np.zeros_like(lhs2)
    o.oracle_point_op(group, 3, O.ptr(lhs2), O.ptr(rhs2), None, O.ptr(exp2), n)
    got2 = M.test_point_op(group, 6, lhs2, rhs2).reshape(n, xb)
    fin2 = exp2[:, half:].any(axis=1)
    assert (got2[:, half:] == exp2[:, half:]).all() and (got2[fin2] == exp2[fin2]).all()
    expd = M.test_point_op(group, 8, lhs2).reshape(n, xb)
    eo = np.zeros_like(lhs2)
    o.oracle_point_op(group, 3, O.ptr(lhs2), O.ptr(lhs2), None, O.ptr(eo), n)  # P + P through the oracle's dadd
    gotd = M.test_point_op(group, 7, lhs2).reshape(n, xb)
    find = eo[:, half:].any(axis=1)
    assert (expd[find] == eo[find]).all() and (gotd[find] == eo[find]).all() and not gotd[~find][:, half:].any()
    # xyzz -> jacobian -> affine, jacobian add / double
    jac = M.test_point_op(group, 4, lhs).reshape(n, jb)
    ej = np.zeros_like(jac)
    o.oracle_point_op(group, 4, O.ptr(lhs), None, None, O.ptr(ej), n)
    fin = lhs[:, half:].any(axis=1)
    assert (jac[fin] == ej[fin]).all()
    aff = M.test_point_op(group, 5, jac).reshape(n, ab)
    ea = np.zeros_like(aff)
    o.oracle_point_op(group, 5, O.ptr(ej), None, None, O.ptr(ea), n)
    assert (aff[fin] == ea[fin]).all() and not aff[~fin].any()
    j2 = np.roll(jac, 3, axis=0).copy()
    j2[:8] = jac[:8]
    for op, rhs_ in ((0, j2), (1, None)):
        exp = np.zeros_like(jac)
        o.oracle_point_op(group, op, O.ptr(jac), O.ptr(rhs_) if rhs_ is not None else None, None, O.ptr(exp), n)
        got = M.test_point_op(group, op, jac, rhs_).reshape(n, jb)
        ga = M.test_point_op(group, 5, got)
        xa = np.zeros(n * ab, dtype=np.uint8)
        o.oracle_point_op(group, 5, O.ptr(exp), None, None, O.ptr(xa), n)
        assert (ga == xa).all()  # compare as group elements (canonical affine)


# ---------------------------------------------------------------- digits (a13, a14, a18, a21 front ends)
@pytest.mark.parametrize("cfgname", ["10", "13", "16", "21"])
def test_digit_decomposition_vs_oracle(M, cfgname):
    n = 512
    ctx = M.MsmContext(1, cfgname, npoints=n)
    oc = O.OracleCtx(1, cfgname, n=n)
    cfg = O.config(cfgname)
    sc = O.gen_scalars(5, n)
    sc[0] = 0
    sc[1] = [1, 0, 0, 0]
    rm1 = O.R_ORDER - 1
    sc[2] = [(rm1 >> (64 * i)) & (2**64 - 1) for i in range(4)]
    B = oc.bucket_set()
    v2i = {int(b): i for i, b in enumerate(B)}
    keys, vals = ctx.digits(0, sc)
    for i in range(n):
        m, b = oc.digits(0, sc[i])
        for j in range(cfg["h"]):
            idx = v2i[int(b[j])]
            slot = i * cfg["h"] + j
            assert keys[i, j] == (idx if idx else 0xFFFFFFFF)
            assert (vals[i, j] & 0x7FFFFFFF) == 3 * slot + abs(int(m[j])) - 1
            assert (vals[i, j] >> 31) == (1 if m[j] < 0 else 0)
    keys, vals = ctx.digits(1, sc)
    trick = cfg["n_exp"] in (13, 14, 16, 17)
    for i in range(n):
        d, cond = oc.digits(1, sc[i])
        for j in range(cfg["h_bgmw"]):
            mag = abs(int(d[j]))
            assert keys[i, j] == (mag if mag else 0xFFFFFFFF)
            sign = (1 if d[j] < 0 else 0) ^ (1 if (trick and cond[0]) else 0)
            if mag:
                assert (vals[i, j] >> 31) == sign
            assert (vals[i, j] & 0x7FFFFFFF) == i * cfg["h_bgmw"] + j
    # Booth windows: sum_t digit_t 2^(t w) == scalar
    w, tiles = ctx.pippenger_window(), ctx.pippenger_tiles()
    assert w == O.config(cfgname)["window"] if n == 1 << cfg["n_exp"] else True
    keys, vals = ctx.digits(2, sc)
    nbw = (1 << (w - 1)) + 1
    ints = O.scalars_to_ints(sc)
    for i in range(n):
        acc = 0
        for t in range(tiles):
            k = int(keys[i, t])
            if k != 0xFFFFFFFF:
                assert k // nbw == t
                mag = k % nbw
                acc += (-mag if (vals[i, t] >> 31) else mag) << (t * w)
        assert acc == ints[i]
    ctx.close()


# ---------------------------------------------------------------- config 10: everything vs the reference driver
@pytest.mark.parametrize("group", [1, 2])
def test_c10_tables_and_four_methods_vs_reference_golden(M, golden, group):
    gd = golden["c10"][str(group)]
    ctx = M.MsmContext(group, "10")
    ctx.init_fix_point_list()
    assert hashlib.sha256(ctx.download(0).tobytes()).hexdigest() == gd["sha256_fix_points"]
    assert hashlib.sha256(ctx.bucket_set().tobytes()).hexdigest() == gd["sha256_bucket_set"]
    ctx.init_pippenger_CHES_q_over_5()
    ctx.init_pippenger_BGMW95()
    assert hashlib.sha256(ctx.download(1).tobytes()).hexdigest() == gd["sha256_table_3nh"]
    assert hashlib.sha256(ctx.download(2).tobytes()).hexdigest() == gd["sha256_table_bgmw95"]
    for seed in (1, 2, 3, 4, 5):
        sc = O.gen_scalars(seed, ctx.n)
        res = [ctx.pippenger_variant_q_over_5_CHES(sc),
               ctx.pippenger_variant_q_over_5_CHES_integral_scalar_conversion(sc),
               ctx.pippenger_variant_BGMW95(sc),
               ctx.pippenger_blst_built_in(sc)]
        for m, r in enumerate(res, 1):
            assert M.affine_serialize(group, r).hex() == gd["msm"][str(seed)], (seed, m)
            assert O.serialize(group, r).hex() == gd["msm"][str(seed)]
    assert ctx.last_launches() > 0
    ctx.close()


@pytest.mark.parametrize("group", [1, 2])
def test_edge_case_scalars(M, group):
    """zero scalars, ones, r-1, all-equal scalars (one giant bucket, P+P inside buckets), mixed."""
    n = 1024
    ctx = M.MsmContext(group, "10")
    ctx.init_fix_point_list()
    ctx.init_pippenger_CHES_q_over_5()
    ctx.init_pippenger_BGMW95()
    rm1 = [((O.R_ORDER - 1) >> (64 * i)) & (2**64 - 1) for i in range(4)]
    cases = {}
    z = np.zeros((n, 4), dtype=np.uint64)
    cases["all_zero"] = z
    one = z.copy(); one[:, 0] = 1
    cases["all_one"] = one
    cases["all_r_minus_1"] = np.array([rm1] * n, dtype=np.uint64)
    same = np.repeat(O.gen_scalars(9, 1), n, axis=0)
    cases["all_equal"] = same
    mixed = O.gen_scalars(10, n); mixed[::3] = 0; mixed[1::7] = rm1
    cases["mixed"] = mixed
    single = z.copy(); single[n - 1] = O.gen_scalars(12, 1)[0]
    cases["single_nonzero_last"] = single
    for name, sc in cases.items():
        exp, _ = O.closed_form(group, sc)
        for method in (1, 2, 3, 4):
            r = ctx.msm(method, sc)
            assert (r == exp).all(), (name, method)
    assert not ctx.msm(1, cases["all_zero"]).any()  # infinity -> all-zero affine (src/e1.c:60-75)
    assert M.affine_serialize(group, ctx.msm(1, cases["all_zero"]))[0] == 0x40
    ctx.close()


def test_last_element_bug_inputs_give_true_sum(M):
    """on inputs where the reference's methods 1-2 drop the last point, the product returns the
    mathematically correct sum (= reference methods 3-4)."""
    n = 1024
    ctx = M.MsmContext(1, "10")
    ctx.init_fix_point_list()
    ctx.init_pippenger_CHES_q_over_5()
    sc = O.gen_scalars(4, n)
    v = O.scalars_to_ints(sc[n - 1:n])[0]
    v &= ~(((1 << 26) - 1) << (13 * 17))
    v |= 1 << (13 * 19)
    v &= (1 << 255) - 1
    sc[n - 1] = [(v >> (64 * i)) & (2**64 - 1) for i in range(4)]
    oc = O.OracleCtx(1, "10", n=n)
    assert oc.hits_bug(sc)
    exp, _ = O.closed_form(1, sc)
    assert (ctx.msm(1, sc) == exp).all() and (ctx.msm(2, sc) == exp).all()
    ctx.close()


@pytest.mark.parametrize("group", [1, 2])
def test_unstructured_points_vs_reference_pippenger_golden(M, golden, group):
    """Caller-supplied (non 2^i G) points incl. an infinity entry: all four methods = compiled reference's
    blst_pNs_mult_pippenger; table build on arbitrary points checked against the oracle's."""
    gd = golden["pippenger_unstructured"][str(group)]
    n = gd["n"]
    pts = np.frombuffer(bytes.fromhex(gd["points"]), dtype=np.uint8).copy()
    sc = np.frombuffer(bytes.fromhex(gd["scalars"]), dtype=np.uint64).reshape(n, 4).copy()
    ctx = M.MsmContext(group, "8", npoints=n)
    ctx.set_points(pts)
    ctx.init_pippenger_CHES_q_over_5()
    ctx.init_pippenger_BGMW95()
    oc = O.OracleCtx(group, "8", n=n, threads=8)
    oc.set_points(pts)
    oc.build_table(0)
    oc.build_table(1)
    assert (ctx.download(1) == oc.table(0)).all()
    assert (ctx.download(2) == oc.table(1)).all()
    for method in (1, 2, 3, 4):
        assert M.affine_serialize(group, ctx.msm(method, sc)).hex() == gd["result"], method
    ctx.close()


def test_ragged_sizes(M):
    """n not a power of two, tiny n."""
    for n in (2, 3, 37, 1000):
        ctx = M.MsmContext(1, "10", npoints=n)
        ctx.init_fix_point_list()
        ctx.init_pippenger_CHES_q_over_5()
        ctx.init_pippenger_BGMW95()
        sc = O.gen_scalars(20 + n, n)
        exp, _ = O.closed_form(1, sc)
        for method in (1, 2, 3, 4):
            assert (ctx.msm(method, sc) == exp).all(), (n, method)
        ctx.close()


def test_state_errors(M):
    ctx = M.MsmContext(1, "10", npoints=16)
    sc = O.gen_scalars(1, 16)
    with pytest.raises(M.MsmB200Error):
        ctx.msm(1, sc)  # no points/table yet
    ctx.init_fix_point_list()
    with pytest.raises(M.MsmB200Error):
        ctx.msm(3, sc)
    with pytest.raises(M.MsmB200Error):
        ctx.msm(7, sc)
    ctx.msm(4, sc)
    ctx.close()


# ---------------------------------------------------------------- blst-named shims (boundary §8b)
@pytest.mark.parametrize("group", [1, 2])
def test_blst_mult_pippenger_shim(M, golden, group):
    gd = golden["pippenger_unstructured"][str(group)]
    n = gd["n"]
    ab, jb = O.AFF_BYTES[group], O.JAC_BYTES[group]
    pts = np.frombuffer(bytes.fromhex(gd["points"]), dtype=np.uint8).copy()
    sc = np.frombuffer(bytes.fromhex(gd["scalars"]), dtype=np.uint8).copy()
    L = M.lib()
    ret = np.zeros(jb, dtype=np.uint8)
    # blst convention: {base, NULL} = contiguous
    pp = (C.c_void_p * 2)(pts.ctypes.data, None)
    sp = (C.c_void_p * 2)(sc.ctypes.data, None)
    getattr(L, "msmb200_blst_p%ds_mult_pippenger" % group)(O.ptr(ret), pp, n, sp, 255, None)
    aff = M.test_point_op(group, 5, ret)
    assert M.affine_serialize(group, aff).hex() == gd["result"]
    # one pointer per element (main_p1.cpp:408-418)
    pp = (C.c_void_p * n)(*[pts.ctypes.data + i * ab for i in range(n)])
    sp = (C.c_void_p * n)(*[sc.ctypes.data + i * 32 for i in range(n)])
    ret2 = np.zeros(jb, dtype=np.uint8)
    getattr(L, "msmb200_blst_p%ds_mult_pippenger" % group)(O.ptr(ret2), pp, n, sp, 255, None)
    assert M.affine_serialize(group, M.test_point_op(group, 5, ret2)).hex() == gd["result"]
    sz = getattr(L, "msmb200_blst_p%ds_mult_pippenger_scratch_sizeof" % group)(n)
    assert sz == O.XYZZ_BYTES[group] << (O.config("8")["window"] - 1 if n == 256 else M.lib().msmb200_pippenger_window_size(n) - 1)


@pytest.mark.parametrize("group", [1, 2])
def test_blst_tile_shims_ches_and_bgmw95(M, group):
    """Drive the literal tile entry points the way main_p1.cpp:208-235,:311-388 does (host pointer arrays into a
    host table, bucket values, signs), table taken from the oracle."""
    n, cfgname = 64, "10"
    cfg = O.config(cfgname)
    ab, jb = O.AFF_BYTES[group], O.JAC_BYTES[group]
    oc = O.OracleCtx(group, cfgname, n=n, threads=4)
    oc.init_fix_points()
    oc.build_table(0)
    oc.build_table(1)
    t3 = oc.table(0)
    tb = oc.table(1)
    sc = O.gen_scalars(31, n)
    exp, _ = O.closed_form(group, sc)
    B = oc.bucket_set()
    q = 1 << cfg["e"]
    v2i = np.zeros(q // 2 + 1, dtype=np.int32)
    v2i[B] = np.arange(len(B), dtype=np.int32)
    h = cfg["h"]
    npts = n * h
    scal = np.zeros(npts + 1, dtype=np.int32)
    signs = np.zeros(npts, dtype=np.uint8)
    ptrs = (C.c_void_p * npts)()
    for i in range(n):
        m, b = oc.digits(0, sc[i])
        for j in range(h):
            k = i * h + j
            scal[k] = b[j]
            signs[k] = 1 if m[j] < 0 else 0
            ptrs[k] = t3.ctypes.data + (3 * k + abs(int(m[j])) - 1) * ab
    ret = np.zeros(jb, dtype=np.uint8)
    Bc = B.copy()
    getattr(M.lib(), "msmb200_blst_p%d_tile_pippenger_d_CHES" % group)(O.ptr(ret), ptrs, npts, O.ptr(scal), O.ptr(signs), None,
                                                                      O.ptr(Bc), O.ptr(v2i), len(B), cfg["d"])
    assert (M.test_point_op(group, 5, ret) == exp).all()
    hb = cfg["h_bgmw"]
    npts = n * hb
    scal = np.zeros(npts, dtype=np.int32)
    signs = np.zeros(npts, dtype=np.uint8)
    ptrs = (C.c_void_p * npts)()
    for i in range(n):
        d, _ = oc.digits(1, sc[i])
        for j in range(hb):
            k = i * hb + j
            scal[k] = abs(int(d[j]))
            signs[k] = 1 if d[j] < 0 else 0
            ptrs[k] = tb.ctypes.data + k * ab
    ret = np.zeros(jb, dtype=np.uint8)
    getattr(M.lib(), "msmb200_blst_p%d_tile_pippenger_BGMW95" % group)(O.ptr(ret), ptrs, npts, O.ptr(scal), O.ptr(signs), None,
                                                                      cfg["e_bgmw"])
    assert (M.test_point_op(group, 5, ret) == exp).all()


@pytest.mark.parametrize("group", [1, 2])
def test_blst_construct_nh_shim_vs_compiled_reference(M, group):
    """msmb200_blst_pN_construct_nh_scalars_nh_points against the compiled reference's blst_pN_construct_nh_scalars_nh_points
    (src/multi_scalar.c:748-775) on the same standard q-ary digits (main_p1.cpp:257-263): bucket values, booth signs and the
    host POINTERS into the caller's 3nh table must be identical, including carries that run through several slots."""
    if not O.has_ref():
        pytest.skip("compiled reference not present")
    cfgname, n = "10", 300
    cfg = O.config(cfgname)
    e, h, q = cfg["e"], cfg["h"], 1 << cfg["e"]
    ab = O.AFF_BYTES[group]
    oc = O.OracleCtx(group, cfgname, n=n)
    H = np.ascontiguousarray(oc.hash_table())  # (q + 1) x {m, b, alpha}, the reference's digit_decomposition layout
    ints = O.scalars_to_ints(O.gen_scalars(41, n))
    ints[0] = 0
    ints[1] = 1
    ints[2] = (1 << (e * (h - 1))) - 1          # every lower digit q - 1: a carry chain through h - 1 slots
    ints[3] = O.R_ORDER - 1 if hasattr(O, "R_ORDER") else ints[3]
    npts = n * h
    digits = np.zeros(npts + 2, dtype=np.int32)  # + 2 slots of slack for the reference's prefetch (main_p1.cpp:254-256)
    for i, s in enumerate(ints):
        for j in range(h):
            digits[i * h + j] = (s >> (e * j)) & (q - 1)
    table = np.zeros((3 * npts, ab), dtype=np.uint8)  # only its address matters
    out = {}
    for name, lib, prefix in (("ref", O.blst_ref(), "blst_p%d" % group), ("cuda", M.lib(), "msmb200_blst_p%d" % group)):
        sc = digits.copy()
        sg = np.full(npts, 7, dtype=np.uint8)
        pp = np.zeros(npts, dtype=np.uint64)
        f = getattr(lib, prefix + "_construct_nh_scalars_nh_points")
        f.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p]
        f.restype = None
        f(O.ptr(sc), O.ptr(sg), O.ptr(pp), npts, O.ptr(table), O.ptr(H))
        out[name] = (sc[:npts], sg, pp)
    assert (out["ref"][0] == out["cuda"][0]).all()
    assert (out["ref"][1] == out["cuda"][1]).all()
    assert (out["ref"][2] == out["cuda"][2]).all()
    assert out["cuda"][1].max() == 1 and (out["cuda"][2] >= table.ctypes.data).all()


@pytest.mark.parametrize("group", [1, 2])
def test_blst_tile_shims_use_the_registered_table(M, group):
    """msmb200_blst_register_table: the host table is mirrored in HBM once, tile calls then upload pointers only (13 bytes per
    entry); a table changed in place is noticed and uploaded again; pointers outside the table take the gather path."""
    n, cfgname = 128, "10"
    cfg = O.config(cfgname)
    ab, jb = O.AFF_BYTES[group], O.JAC_BYTES[group]
    hb = cfg["h_bgmw"]
    oc = O.OracleCtx(group, cfgname, n=n, threads=4)
    oc.init_fix_points()
    oc.build_table(1)
    tb = np.ascontiguousarray(oc.table(1)).reshape(-1, ab)
    assert M.lib().msmb200_blst_register_table(group, O.ptr(tb), tb.shape[0]) == 0
    tile = getattr(M.lib(), "msmb200_blst_p%d_tile_pippenger_BGMW95" % group)

    def run(table, sc, rows=None):
        rows = range(n) if rows is None else rows
        npts = len(rows) * hb
        scal = np.zeros(npts, dtype=np.int32)
        signs = np.zeros(npts, dtype=np.uint8)
        ptrs = (C.c_void_p * npts)()
        for t, i in enumerate(rows):
            d, _ = oc.digits(1, sc[i])
            for j in range(hb):
                scal[t * hb + j] = abs(int(d[j]))
                signs[t * hb + j] = 1 if d[j] < 0 else 0
                ptrs[t * hb + j] = table.ctypes.data + (i * hb + j) * ab
        ret = np.zeros(jb, dtype=np.uint8)
        tile(O.ptr(ret), ptrs, npts, O.ptr(scal), O.ptr(signs), None, cfg["e_bgmw"])
        return M.test_point_op(group, 5, ret)

    sc = O.gen_scalars(61, n)
    exp, _ = O.closed_form(group, sc)
    assert (run(tb, sc) == exp).all()
    assert (run(tb, sc) == exp).all()           # second call: table already resident
    # the same buffer now holds the table of OTHER points (rows 3.. then 0..2): the mirror must follow
    perm = np.roll(np.arange(n), -3)
    tb[:] = tb.reshape(n, hb, ab)[perm].reshape(-1, ab)
    sc2 = sc[perm]
    assert (run(tb, sc2) == exp).all()
    # pointers into a different buffer (a copy): learned / gathered, still right
    other = tb.copy()
    assert (run(other, sc2) == exp).all()
    assert M.lib().msmb200_blst_last_call_ms(group) > 0


def _dropin_driver(group, suffix=""):
    import os

    path = os.path.join(O.REF_DIR, "refdrv_p%d_dropin%s.so" % (group, suffix))
    if not os.path.exists(path):
        pytest.skip("%s not built (needs /root/reference at build time)" % os.path.basename(path))
    lib = C.CDLL(path)
    lib.refdrv_msm.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.refdrv_last_ms.restype = C.c_double
    fd = os.dup(1)  # the driver's init reports progress on std::cout
    devnull = os.open(os.devnull, os.O_WRONLY)
    os.dup2(devnull, 1)
    try:
        lib.refdrv_init()
    finally:
        os.dup2(fd, 1)
        os.close(fd)
        os.close(devnull)
    return lib


@pytest.mark.parametrize("group", [1, 2])
def test_reference_driver_glue_on_the_cuda_shims_vs_golden(M, golden, group):
    """The UNMODIFIED main_p{1,2}.cpp (config 10) as a library whose five MSM imports are re-pointed at libmsm_b200.so
    (oracle/Makefile: refdrv_pN_dropin.so): its own init, digit conversions and pointer arrays, the CUDA shims underneath,
    on the SEEDED scalars of the golden vectors - each of its four methods must print the reference's own result."""
    lib = _dropin_driver(group)
    gd = golden["c10"][str(group)]
    for seed in (1, 2, 3):
        sc = O.gen_scalars(seed, 1 << 10)
        for method in (1, 2, 3, 4):
            ob = np.zeros(O.SER_BYTES[group], dtype=np.uint8)
            assert lib.refdrv_msm(method, O.ptr(sc), None, O.ptr(ob)) == 0
            assert ob.tobytes().hex() == gd["msm"][str(seed)], (seed, method)


@pytest.mark.parametrize("cfgname", ["13", "16"])
def test_reference_driver_glue_other_configs_and_shim_cost(M, cfgname):
    """The drop-in driver at configuration 13 (always) and 16 (MSMB200_SLOW_TESTS=1: its CPU table build takes about a
    minute): results against the oracle's closed form; the time one tile-shim call costs (pointers translated on the device
    against the mirrored table) next to the context API's host-to-host time for the same MSM."""
    import os
    import time

    if cfgname == "16" and not os.environ.get("MSMB200_SLOW_TESTS"):
        pytest.skip("set MSMB200_SLOW_TESTS=1 (the reference's CPU table build at n = 2^16 takes about a minute)")
    lib = _dropin_driver(1, "_c" + cfgname)
    n = 1 << int(cfgname)
    ctx = M.MsmContext(1, cfgname)
    ctx.init_fix_point_list()
    ctx.init_pippenger_CHES_q_over_5()
    ctx.init_pippenger_BGMW95()
    report = {}
    for seed in (11, 12):
        sc = O.gen_scalars(seed, n)
        exp = O.serialize(1, O.closed_form(1, sc)[0])
        for method in (1, 2, 3, 4):
            ob = np.zeros(O.SER_BYTES[1], dtype=np.uint8)
            assert lib.refdrv_msm(method, O.ptr(sc), None, O.ptr(ob)) == 0
            assert ob.tobytes() == exp, (seed, method)
            t0 = time.perf_counter()
            got = ctx.msm(method, sc)
            api_ms = (time.perf_counter() - t0) * 1e3
            assert O.serialize(1, got) == exp
            report[method] = (lib.refdrv_last_ms(), M.lib().msmb200_blst_last_call_ms(1), api_ms)
    for method, (drv, shim, api) in report.items():
        print("config %s method %d: driver method %.2f ms (reference glue on the host included), last shim call %.2f ms, "
              "context API host-to-host %.2f ms" % (cfgname, method, drv, shim, api))
    ctx.close()


# ---------------------------------------------------------------- multi-GPU path emulated on one device
@pytest.mark.parametrize("group", [1, 2])
def test_sharded_partials_sum_to_full_result(M, group):
    """G = 3 shards as three contexts on one GPU (B200_PROFILING.md: emulate ranks in one process): each yields a
    Jacobian partial in device memory; sum_partials == single-context result == closed form."""
    import torch

    from msm_blst_b200 import distributed as D

    n, world = 1000, 3
    sc = O.gen_scalars(41, n)
    exp, _ = O.closed_form(group, sc)
    jb = O.JAC_BYTES[group]
    partials = torch.zeros((world, jb), dtype=torch.uint8, device="cuda")
    ctxs = []
    for r in range(world):
        lo, hi = D.shard_range(n, r, world)
        ctx = M.MsmContext(group, "10", npoints=hi - lo, first=lo)
        ctx.init_fix_point_list()
        ctx.init_pippenger_CHES_q_over_5()
        d_sc = torch.from_numpy(sc[lo:hi].view(np.uint8).copy()).cuda()
        ctx.msm_partial_device(1, d_sc.data_ptr(), partials[r].data_ptr())
        ctxs.append((ctx, d_sc))
    torch.cuda.synchronize()
    got = ctxs[0][0].sum_partials_device(partials.data_ptr(), world)
    assert (got == exp).all()
    host = partials.cpu().numpy().copy()
    chk = np.zeros(O.AFF_BYTES[group], dtype=np.uint8)
    O.oracle().oracle_sum_partials(group, O.ptr(host), world, O.ptr(chk))
    assert (chk == exp).all()
    for ctx, _ in ctxs:
        ctx.close()


# ---------------------------------------------------------------- the unmodified reference driver on the CUDA shims
@pytest.mark.parametrize("group", [1, 2])
def test_reference_driver_runs_on_the_cuda_shims(M, group):
    """oracle/_ref/main_p{1,2}_dropin = the reference's own main_p{1,2}.cpp (config 10), compiled in place with its
    MSM imports re-pointed at libmsm_b200.so. Its four methods run on fresh OpenSSL-random
    scalars; they use different tables / digit systems, so identical printed points mean the shims are drop-ins."""
    import os
    import re
    import subprocess

    exe = os.path.join(O.REF_DIR, "main_p%d_dropin" % group)
    if not os.path.exists(exe):
        pytest.skip("drop-in driver not built (needs /root/reference at build time)")
    out = subprocess.run("ulimit -s unlimited 2>/dev/null; exec %s" % exe, shell=True, capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stderr[-2000:]
    text = out.stdout
    blocks = re.split(r"\n\d\. ", text)[1:]
    assert len(blocks) >= 4, text[-3000:]
    pts = []
    for b in blocks[:4]:
        coords = re.findall(r"0x[0-9a-f ]{90,}", b)
        assert len(coords) >= (2 if group == 1 else 4), b[:500]
        pts.append(tuple(coords[: (2 if group == 1 else 4)]))
    assert pts[0] == pts[1] == pts[2] == pts[3], pts
    assert "0x0000000000000000 0000000000000000 0000000000000000 0000000000000000 0000000000000000 0000000000000000" not in pts[0][0]


# ---------------------------------------------------------------- both bucket accumulators
@pytest.mark.parametrize("group", [1, 2])
@pytest.mark.parametrize("mode", [1, 2])
def test_accumulator_modes_agree(M, group, mode):
    """mode 1 = XYZZ mixed additions per work item (xyzz_dadd_affine loop), mode 2 = batch-affine pairwise rounds
    with one shared inversion per batch (bulk_addition.c analogue). Same bytes, including the inputs that force
    P + P, P - P and infinity inside a batch (all-equal scalars, zero scalars, infinity table entries)."""
    n = 1024
    ctx = M.MsmContext(group, "10")
    ctx.set_accumulator(mode)
    ctx.init_fix_point_list()
    ctx.init_pippenger_CHES_q_over_5()
    ctx.init_pippenger_BGMW95()
    rm1 = [((O.R_ORDER - 1) >> (64 * i)) & (2**64 - 1) for i in range(4)]
    cases = [O.gen_scalars(50 + mode, n), np.repeat(O.gen_scalars(9, 1), n, axis=0), np.zeros((n, 4), dtype=np.uint64),
             np.array([rm1] * n, dtype=np.uint64)]
    mixed = O.gen_scalars(10, n)
    mixed[::3] = 0
    cases.append(mixed)
    for sc in cases:
        exp, _ = O.closed_form(group, sc)
        for method in (1, 2, 3, 4):
            assert (ctx.msm(method, sc) == exp).all(), (mode, method)
    ctx.close()
    # unstructured points with an infinity entry in the table
    import json, os
    gd = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "golden.json")))["pippenger_unstructured"][str(group)]
    pts = np.frombuffer(bytes.fromhex(gd["points"]), dtype=np.uint8).copy()
    sc = np.frombuffer(bytes.fromhex(gd["scalars"]), dtype=np.uint64).reshape(gd["n"], 4).copy()
    ctx = M.MsmContext(group, "8", npoints=gd["n"])
    ctx.set_accumulator(mode)
    ctx.set_points(pts)
    ctx.init_pippenger_CHES_q_over_5()
    for method in (1, 4):
        assert M.affine_serialize(group, ctx.msm(method, sc)).hex() == gd["result"]
    ctx.close()


@pytest.mark.parametrize("group", [1, 2])
@pytest.mark.parametrize("reducer", [1, 2])
@pytest.mark.parametrize("accum", [1, 2])
def test_reducer_modes_agree(M, group, reducer, accum):
    """Both bucket reductions (1 = chunked running sums with the reference's tmp_d[] gap accumulators, 2 = digit
    splitting into per-digit / per-bit lists) under both accumulators, all four methods, on inputs that leave most
    buckets empty (zero / equal scalars) as well as on dense ones; config 13 has the r - a trick in BGMW95."""
    for cfgname in ("10", "13"):
        ctx = M.MsmContext(group, cfgname)
        ctx.set_reducer(reducer)
        ctx.set_accumulator(accum)
        ctx.init_fix_point_list()
        ctx.init_pippenger_CHES_q_over_5()
        ctx.init_pippenger_BGMW95()
        n = ctx.n
        rm1 = [((O.R_ORDER - 1) >> (64 * i)) & (2**64 - 1) for i in range(4)]
        cases = [O.gen_scalars(70 + reducer, n), np.repeat(O.gen_scalars(9, 1), n, axis=0), np.zeros((n, 4), dtype=np.uint64),
                 np.array([rm1] * n, dtype=np.uint64)]
        one = np.zeros((n, 4), dtype=np.uint64)
        one[n // 2, 0] = 1  # a single non-empty bucket
        cases.append(one)
        for sc in cases:
            exp, _ = O.closed_form(group, sc)
            for method in (1, 2, 3, 4):
                assert (ctx.msm(method, sc) == exp).all(), (cfgname, reducer, accum, method)
        ctx.close()


def test_batch_affine_full_size_known_answer(M, golden):
    ctx = M.MsmContext(1, "16")
    ctx.set_accumulator(2)
    ctx.init_fix_point_list()
    ctx.init_pippenger_CHES_q_over_5()
    sc = O.gen_scalars(1, ctx.n)
    assert M.affine_serialize(1, ctx.msm(1, sc)).hex() == golden["kat_appc"]["g1_n16"]
    ctx.close()


@pytest.mark.parametrize("group,cfgname,reps", [(1, "16", 12), (2, "13", 12), (1, "20", 3)])
def test_batch_affine_rounds_repeatable_on_fresh_scalars(M, group, cfgname, reps):
    """The batch-affine rounds stage their operands through shared memory with cp.async, fill each other's stages in round 0
    (warp-cooperative gather) and take their batches from an atomic counter, so the work split differs from run to run: every
    repetition, on a fresh scalar set and back to back without synchronisation in between, must still give the closed form;
    the XYZZ accumulator is the cross-check on the first set. Also covers a table in the packed layout (pointers of a caller)
    next to the context's padded one: method 4 gathers from the packed fixed points."""
    ctx = M.MsmContext(group, cfgname)
    ctx.init_fix_point_list()
    ctx.init_pippenger_CHES_q_over_5()
    ctx.set_accumulator(2)
    for rep in range(reps):
        sc = O.gen_scalars(900 + 7 * rep + group, ctx.n)
        exp, _ = O.closed_form(group, sc)
        got = ctx.msm(1, sc)
        assert ctx.last_accumulator() == 2
        assert (got == exp).all(), rep
        if rep == 0:
            assert (ctx.msm(4, sc) == exp).all()
            ctx.set_accumulator(1)
            assert (ctx.msm(1, sc) == exp).all() and ctx.last_accumulator() == 1
            ctx.set_accumulator(2)
    ctx.close()


@pytest.mark.parametrize("group", [1, 2])
def test_packed_and_padded_tables_agree(M, group, monkeypatch):
    """The context's own tables are padded to whole 128-byte lines; MSMB200_PACKED_TABLES keeps the reference's packed layout
    (the layout caller-owned tables have). Both accumulators must give the same result on both, and the download is the
    packed reference layout either way."""
    cfgname = "13"
    sc = O.gen_scalars(77 + group, 1 << 13)
    exp, _ = O.closed_form(group, sc)
    tables = []
    for packed in (False, True):
        if packed:
            monkeypatch.setenv("MSMB200_PACKED_TABLES", "1")
        ctx = M.MsmContext(group, cfgname)
        ctx.init_fix_point_list()
        ctx.init_pippenger_CHES_q_over_5()
        ctx.init_pippenger_BGMW95()
        for accum in (1, 2):
            ctx.set_accumulator(accum)
            for method in (1, 2, 3):
                assert (ctx.msm(method, sc) == exp).all(), (packed, accum, method)
        tables.append((ctx.download(1).copy(), ctx.download(2).copy()))
        ctx.close()
    assert (tables[0][0] == tables[1][0]).all() and (tables[0][1] == tables[1][1]).all()


@pytest.mark.parametrize("field", [1, 2])
def test_warp_batch_inversion_vs_plain_inverse(M, field):
    """Montgomery's trick across the warp (one inversion per 32 lanes, used by the batch-affine rounds) returns the
    same bytes as the per-element inverse (reference reciprocal_fp / reciprocal_fp2, src/recip.c:58-114)."""
    rng = np.random.default_rng(3)
    n = 3000 + 37
    a = rng.integers(0, 2**64, size=(n, 6 * field), dtype=np.uint64)
    a[:, 5::6] &= np.uint64(0x0FFFFFFFFFFFFFFF)
    one = np.frombuffer(bytes.fromhex("fdff02000000097602000cc40b00f4ebba58c7535798485f455752705358ce776dec56a2971a075c93e480fac35ef615"), dtype=np.uint64)
    idx = rng.integers(0, n, size=n // 3)
    a[idx, :6] = one
    if field == 2:
        a[idx, 6:] = 0
    exp = np.empty_like(a)
    (O.oracle().oracle_fp_op if field == 1 else O.oracle().oracle_fp2_op)(6, O.ptr(a), None, O.ptr(exp), n)
    assert (M.test_field_op(field, 7, a) == exp).all()
    assert (M.test_field_op(field, 6, a) == exp).all()


@pytest.mark.parametrize("group", [1, 2])
@pytest.mark.parametrize("mode", ["1", "2"])
def test_blst_points_add_shim_vs_compiled_reference(M, golden, group, mode, monkeypatch):
    """msmb200_blst_pNs_add (sum of affine points; the reference's bulk_addition.c consumer) against the
    compiled reference's blst_pNs_add when oracle/_ref is present, else against the oracle; both accumulators."""
    monkeypatch.setenv("MSMB200_ACCUM", mode)
    gd = golden["pippenger_unstructured"][str(group)]
    n = gd["n"]
    ab, jb = O.AFF_BYTES[group], O.JAC_BYTES[group]
    pts = np.frombuffer(bytes.fromhex(gd["points"]), dtype=np.uint8).copy()
    # duplicate a few points so that P + P shows up in the tree, and add P, -P neighbours
    pts2 = np.concatenate([pts, pts[: 10 * ab], pts[20 * ab: 21 * ab]])
    n2 = n + 11
    pp = (C.c_void_p * 2)(pts2.ctypes.data, None)
    ret = np.zeros(jb, dtype=np.uint8)
    getattr(M.lib(), "msmb200_blst_p%ds_add" % group)(O.ptr(ret), pp, n2)
    got = M.test_point_op(group, 5, ret)
    # expected: sum with all-one scalars through the oracle (naive double-and-add MSM)
    ones = np.zeros((n2, 4), dtype=np.uint64)
    ones[:, 0] = 1
    exp = np.zeros(ab, dtype=np.uint8)
    O.oracle().oracle_naive_msm(group, O.ptr(pts2), O.ptr(ones), n2, O.ptr(exp))
    assert (got == exp).all()
    if O.has_ref():
        b = O.blst_ref()
        rj = np.zeros(jb, dtype=np.uint8)
        getattr(b, "blst_p%ds_add" % group)(O.ptr(rj), pp, C.c_size_t(n2))
        ra = np.zeros(ab, dtype=np.uint8)
        getattr(b, "blst_p%d_to_affine" % group)(O.ptr(ra), O.ptr(rj))
        assert (got == ra).all()


@pytest.mark.parametrize("group", [1, 2])
@pytest.mark.parametrize("wbits", [4, 7])
def test_blst_mult_wbits_shims_vs_compiled_reference(M, golden, group, wbits):
    """msmb200_blst_pNs_mult_wbits_precompute / _mult_wbits (the library's fixed-window table MSM, SURVEY §8f-1;
    src/multi_scalar.c:81-261): the table is byte-identical to the compiled reference's, and the MSM over it equals
    the oracle's naive MSM and the reference's own blst_pNs_mult_wbits after to_affine. nbits = 255 and 64."""
    gd = golden["pippenger_unstructured"][str(group)]
    ab, jb = O.AFF_BYTES[group], O.JAC_BYTES[group]
    pts = np.frombuffer(bytes.fromhex(gd["points"]), dtype=np.uint8).reshape(gd["n"], ab)
    pts = np.ascontiguousarray(pts[pts.any(axis=1)])  # the reference's precompute is undefined for infinity inputs
    n = pts.shape[0]
    nwin = 1 << (wbits - 1)
    L = M.lib()
    size_fn = getattr(L, "msmb200_blst_p%ds_mult_wbits_precompute_sizeof" % group)
    size_fn.restype = C.c_size_t
    assert size_fn(C.c_size_t(wbits), C.c_size_t(n)) == n * nwin * ab
    pp = (C.c_void_p * 2)(pts.ctypes.data, None)
    table = np.zeros((n * nwin, ab), dtype=np.uint8)
    getattr(L, "msmb200_blst_p%ds_mult_wbits_precompute" % group)(O.ptr(table), C.c_size_t(wbits), pp, C.c_size_t(n))
    # row entry k of point i is (k + 1) * P_i: spot-check through the oracle's naive MSM
    rng = np.random.default_rng(wbits)
    for i, k in zip(rng.integers(0, n, size=6), rng.integers(0, nwin, size=6)):
        sc1 = np.zeros((1, 4), dtype=np.uint64)
        sc1[0, 0] = k + 1
        e = np.zeros(ab, dtype=np.uint8)
        O.oracle().oracle_naive_msm(group, O.ptr(np.ascontiguousarray(pts[i])), O.ptr(sc1), 1, O.ptr(e))
        assert (table[i * nwin + k] == e).all(), (i, k)
    assert (table[::nwin] == pts).all()
    ref = O.blst_ref() if O.has_ref() else None
    if ref is not None:
        rt = np.zeros_like(table)
        getattr(ref, "blst_p%ds_mult_wbits_precompute" % group)(O.ptr(rt), C.c_size_t(wbits), pp, C.c_size_t(n))
        assert (rt == table).all()
    sc = O.gen_scalars(77 + wbits, n)
    sc[3] = 0
    for nbits in (255, 64):
        scb = np.ascontiguousarray(sc.view(np.uint8).reshape(n, 32)[:, : (nbits + 7) // 8])
        sp = (C.c_void_p * 2)(scb.ctypes.data, None)
        ret = np.zeros(jb, dtype=np.uint8)
        getattr(L, "msmb200_blst_p%ds_mult_wbits" % group)(O.ptr(ret), O.ptr(table), C.c_size_t(wbits), C.c_size_t(n), sp, C.c_size_t(nbits), None)
        got = M.test_point_op(group, 5, ret)
        sct = sc.copy()
        if nbits == 64:
            sct[:, 1:] = 0
        exp = np.zeros(ab, dtype=np.uint8)
        O.oracle().oracle_naive_msm(group, O.ptr(pts), O.ptr(sct), n, O.ptr(exp))
        assert (got == exp).all(), nbits
        if ref is not None:
            rj = np.zeros(jb, dtype=np.uint8)
            getattr(ref, "blst_p%ds_mult_wbits" % group)(O.ptr(rj), O.ptr(table), C.c_size_t(wbits), C.c_size_t(n), sp, C.c_size_t(nbits), None)
            ra = np.zeros(ab, dtype=np.uint8)
            getattr(ref, "blst_p%d_to_affine" % group)(O.ptr(ra), O.ptr(rj))
            assert (got == ra).all(), nbits


def _booth_tile_digit(s, nbits, bit0, window):
    """Signed digit of scalar s for the tile (bit0, window): restatement of src/multi_scalar.c:587-600 + :395-400 +
    booth_encode (src/ec_mult.h:45-56)."""
    if bit0 + window > nbits:
        wbits = nbits - bit0
        cbits = wbits + 1
    else:
        wbits = cbits = window
    wmask = (1 << (wbits + 1)) - 1
    z = 1 if bit0 == 0 else 0
    b0, wb = bit0 - (z ^ 1), wbits + (z ^ 1)
    wval = (((s >> b0) & ((1 << wb) - 1)) << z) & wmask
    sign = (wval >> cbits) & 1
    v = (wval + 1) >> 1
    return v - (1 << cbits) if sign else v


@pytest.mark.parametrize("group", [1, 2])
def test_blst_tile_pippenger_shim(M, golden, group):
    """msmb200_blst_pNs_tile_pippenger (one window, the grid entry point of the upstream bindings):
    each tile equals sum_i digit_i * P_i with the reference's Booth digits (oracle naive MSM on digit mod r), the tiles
    of a scalar recombine to the scalar, and the bytes equal the compiled reference's tile when oracle/_ref is present."""
    gd = golden["pippenger_unstructured"][str(group)]
    ab, jb = O.AFF_BYTES[group], O.JAC_BYTES[group]
    pts = np.frombuffer(bytes.fromhex(gd["points"]), dtype=np.uint8).copy()
    n = gd["n"]
    sc = O.gen_scalars(91, n)
    ints = O.scalars_to_ints(sc)
    nbits, window = 255, 7
    assert all(sum(_booth_tile_digit(s, nbits, b, window) << b for b in range(0, nbits + 1, window)) == s for s in ints[:50])
    scb = np.ascontiguousarray(sc.view(np.uint8).reshape(n, 32))
    pp = (C.c_void_p * 2)(pts.ctypes.data, None)
    sp = (C.c_void_p * 2)(scb.ctypes.data, None)
    fn = getattr(M.lib(), "msmb200_blst_p%ds_tile_pippenger" % group)
    ref = O.blst_ref() if O.has_ref() else None
    for bit0, w in ((0, 7), (7, 7), (119, 7), (252, 7), (245, 10), (13, 5)):
        ret = np.zeros(jb, dtype=np.uint8)
        fn(O.ptr(ret), pp, C.c_size_t(n), sp, C.c_size_t(nbits), None, C.c_size_t(bit0), C.c_size_t(w))
        got = M.test_point_op(group, 5, ret)
        dig = np.zeros((n, 4), dtype=np.uint64)
        for i, s in enumerate(ints):
            d = _booth_tile_digit(s, nbits, bit0, w) % O.R_ORDER
            dig[i] = [(d >> (64 * k)) & (2**64 - 1) for k in range(4)]
        exp = np.zeros(ab, dtype=np.uint8)
        O.oracle().oracle_naive_msm(group, O.ptr(pts), O.ptr(dig), n, O.ptr(exp))
        assert (got == exp).all(), (bit0, w)
        if ref is not None:
            scratch = np.zeros(O.XYZZ_BYTES[group] << (w - 1), dtype=np.uint8)
            rj = np.zeros(jb, dtype=np.uint8)
            getattr(ref, "blst_p%ds_tile_pippenger" % group)(O.ptr(rj), pp, C.c_size_t(n), sp, C.c_size_t(nbits), O.ptr(scratch), C.c_size_t(bit0), C.c_size_t(w))
            ra = np.zeros(ab, dtype=np.uint8)
            getattr(ref, "blst_p%d_to_affine" % group)(O.ptr(ra), O.ptr(rj))
            assert (got == ra).all(), (bit0, w)


@pytest.mark.parametrize("group", [1, 2])
def test_blst_points_to_affine_shim(M, group):
    """msmb200_blst_pNs_to_affine (batched normalisation, src/multi_scalar.c:17-59): same bytes as the
    per-point blst_pN_to_affine of the oracle and, when present, as the compiled reference's batched call; infinity inputs
    (Z = 0) give (0, 0); batch sizes that are not a multiple of the 3-point inversion groups."""
    ab, jb, xb = O.AFF_BYTES[group], O.JAC_BYTES[group], O.XYZZ_BYTES[group]
    oc = O.OracleCtx(group, "10", n=64)
    oc.init_fix_points()
    pts = oc.points().reshape(64, ab)
    rng = np.random.default_rng(11)
    n = 1000 + 1
    # Jacobian points with non-trivial Z: sums of XYZZ accumulations converted with op 4
    acc = np.zeros((n, xb), dtype=np.uint8)
    for rounds in range(3):
        acc = M.test_point_op(group, 2, acc, pts[rng.integers(0, 64, size=n)].copy(), rng.integers(0, 2, size=n).astype(np.uint8)).reshape(n, xb)
    jac = M.test_point_op(group, 4, acc).reshape(n, jb).copy()
    jac[5] = 0
    jac[n - 1] = 0  # infinity in the ragged last group
    for count in (n, n - 1, n - 2, 1, 2):
        sub = np.ascontiguousarray(jac[:count])
        pp = (C.c_void_p * 2)(sub.ctypes.data, None)
        dst = np.zeros((count, ab), dtype=np.uint8)
        getattr(M.lib(), "msmb200_blst_p%ds_to_affine" % group)(O.ptr(dst), pp, C.c_size_t(count))
        exp = np.zeros_like(dst)
        O.oracle().oracle_point_op(group, 5, O.ptr(sub), None, None, O.ptr(exp), count)
        assert (dst == exp).all(), count
        if count > 5:
            assert not dst[5].any()
    if O.has_ref():  # the reference's batched call is undefined for Z = 0 entries: compare on a finite range
        sub = np.ascontiguousarray(jac[10:510])
        pp = (C.c_void_p * 2)(sub.ctypes.data, None)
        dst = np.zeros((500, ab), dtype=np.uint8)
        getattr(M.lib(), "msmb200_blst_p%ds_to_affine" % group)(O.ptr(dst), pp, C.c_size_t(500))
        rd = np.zeros_like(dst)
        getattr(O.blst_ref(), "blst_p%ds_to_affine" % group)(O.ptr(rd), pp, C.c_size_t(500))
        assert (rd == dst).all()


# ---------------------------------------------------------------- table persistence
@pytest.mark.parametrize("group", [1, 2])
def test_table_save_load_round_trip_and_serialized_bytes(M, group, tmp_path):
    """msmb200_table_save / _load: format 0 is the in-memory blst_pN_affine layout, format 1 is blst_pN_affine_serialize
    of every entry (checked against the compiled reference when present and against the oracle-pinned
    msmb200_affine_serialize); a fresh context loads points and both tables from disk and reproduces the MSM; a
    corrupted entry and a file written for another configuration are rejected."""
    ab = O.AFF_BYTES[group]
    HDR = 80  # magic, version, group, format, which, configuration, npoints, entries, digest of the fixed points
    ctx = M.MsmContext(group, "10")
    ctx.init_fix_point_list()
    ctx.init_pippenger_CHES_q_over_5()
    ctx.init_pippenger_BGMW95()
    sc = O.gen_scalars(5, ctx.n)
    exp, _ = O.closed_form(group, sc)
    files = {}
    for which in (0, 1, 2):
        for fmt in (0, 1):
            files[(which, fmt)] = str(tmp_path / ("t%d_%d.bin" % (which, fmt)))
            ctx.table_save(which, files[(which, fmt)], fmt)
    tbl = ctx.download(1)
    raw = np.fromfile(files[(1, 0)], dtype=np.uint8)
    ser = np.fromfile(files[(1, 1)], dtype=np.uint8)
    assert raw.size == HDR + tbl.size and ser.size == raw.size and bytes(raw[:8]) == b"MSMB200T"
    assert (raw[HDR:] == tbl.reshape(-1)).all()
    entries = tbl.reshape(-1, ab)
    ser = ser[HDR:].reshape(-1, ab)
    ref = O.blst_ref() if O.has_ref() else None
    for k in (0, 1, 2, 777, entries.shape[0] - 1):
        assert bytes(ser[k]) == bytes(M.affine_serialize(group, entries[k]))
        if ref is not None:
            out = np.zeros(ab, dtype=np.uint8)
            getattr(ref, "blst_p%d_affine_serialize" % group)(O.ptr(out), O.ptr(np.ascontiguousarray(entries[k])))
            assert (out == ser[k]).all()
    ctx.close()
    for fmt in (0, 1):
        c2 = M.MsmContext(group, "10")
        with pytest.raises(M.MsmB200Error):
            c2.msm(1, sc)  # nothing loaded yet
        for which in (0, 1, 2):
            c2.table_load(which, files[(which, fmt)])
        assert (c2.download(1) == tbl).all()
        for method in (1, 2, 3, 4):
            assert (c2.msm(method, sc) == exp).all(), (fmt, method)
        c2.close()
    # a flipped byte inside one entry -> not a curve point (or out of range): rejected
    for fmt in (0, 1):
        bad = np.fromfile(files[(1, fmt)], dtype=np.uint8)
        bad[HDR + 5 * ab + 40] ^= 1
        badpath = str(tmp_path / ("bad%d.bin" % fmt))
        bad.tofile(badpath)
        c3 = M.MsmContext(group, "10")
        c3.table_load(0, files[(0, fmt)])
        with pytest.raises(M.MsmB200Error):
            c3.table_load(1, badpath)
        with pytest.raises(M.MsmB200Error):
            c3.msm(1, sc)  # the rejected table is not usable
        c3.close()
    # written for config 10 (e = 13, h = 20): a config-11 context (e = 14, h = 19) refuses it, as does the other slot
    c4 = M.MsmContext(group, "11", npoints=1024)
    with pytest.raises(M.MsmB200Error):
        c4.table_load(1, files[(1, 0)])
    with pytest.raises(M.MsmB200Error):
        c4.table_load(2, files[(1, 0)])
    c4.close()
    # same group / size / configuration but OTHER fixed points (the shard that starts at point 1024): the table is refused,
    # and so is a table offered before any points are present
    c5 = M.MsmContext(group, "10", first=1024)
    with pytest.raises(M.MsmB200Error):
        c5.table_load(1, files[(1, 0)])
    c5.init_fix_point_list()
    with pytest.raises(M.MsmB200Error):
        c5.table_load(1, files[(1, 0)])
    with pytest.raises(M.MsmB200Error):
        c5.table_load(2, files[(2, 1)])
    c5.close()


@pytest.mark.parametrize("group", [1, 2])
def test_sharded_per_bit_sums_combine_to_full_result(M, group):
    """The multi-GPU leg that runs ONE Horner pass and ONE inversion for the whole job: G = 3 shards as contexts on one
    GPU, each stops after its bucket reduction and writes its per-bit XYZZ sums (msmb200_msm_bits_device); the gathered
    sums are added entry by entry over the ranks, shifted and normalised once (msmb200_combine_bits_device). All four
    methods (blst-Pippenger has 32 windows at n = 2^10), result == closed form; also through distributed.msm_sharded's
    single-rank path."""
    import torch

    from msm_blst_b200 import distributed as D

    n, world = 1000, 3
    sc = O.gen_scalars(44, n)
    exp, _ = O.closed_form(group, sc)
    ctxs = []
    for r in range(world):
        lo, hi = D.shard_range(n, r, world)
        ctx = M.MsmContext(group, "10", npoints=hi - lo, first=lo)
        ctx.init_fix_point_list()
        ctx.init_pippenger_CHES_q_over_5()
        ctx.init_pippenger_BGMW95()
        ctxs.append((ctx, torch.from_numpy(sc[lo:hi].view(np.uint8).copy()).cuda()))
    for method in (1, 2, 3, 4):
        lays = {ctx.msm_bits_layout(method) for ctx, _ in ctxs}
        if method == 4 and lays != {None} and len(lays) > 1:
            continue  # shards of different sizes choose different Pippenger windows: nothing to combine entry by entry
        if lays == {None}:
            continue
        assert len(lays) == 1
        lay = lays.pop()
        per = lay[0] * lay[1] * O.XYZZ_BYTES[group] if hasattr(O, "XYZZ_BYTES") else lay[0] * lay[1] * (192 if group == 1 else 384)
        gathered = torch.zeros((world, per), dtype=torch.uint8, device="cuda")
        for r, (ctx, d_sc) in enumerate(ctxs):
            ctx.msm_bits_device(method, d_sc.data_ptr(), gathered[r].data_ptr())
        torch.cuda.synchronize()
        got = ctxs[1][0].combine_bits_device(gathered.data_ptr(), world, lay)
        assert (got == exp).all(), (method, lay)
    # one rank: msm_sharded without a process group is not possible; the combine with world = 1 must equal msm_device
    ctx, d_sc = ctxs[0]
    lay = ctx.msm_bits_layout(1)
    buf = torch.zeros(lay[0] * lay[1] * (192 if group == 1 else 384), dtype=torch.uint8, device="cuda")
    ctx.msm_bits_device(1, d_sc.data_ptr(), buf.data_ptr())
    torch.cuda.synchronize()
    assert (ctx.combine_bits_device(buf.data_ptr(), 1, lay) == ctx.msm_device(1, d_sc.data_ptr())).all()
    for ctx, _ in ctxs:
        ctx.close()


# ---------------------------------------------------------------- bucket-range sharding (tables replicated)
@pytest.mark.parametrize("group", [1, 2])
def test_bucket_range_shards_sum_to_full_result(M, group):
    """world = 3 contexts on one GPU, each holding ALL points and owning one third of the bucket-reduction chunks:
    the Jacobian partials sum to the single-context result for every method."""
    import torch

    n, world = 1024, 3
    sc = O.gen_scalars(43, n)
    exp, _ = O.closed_form(group, sc)
    jb = O.JAC_BYTES[group]
    d_sc = torch.from_numpy(sc.view(np.uint8).copy()).cuda()
    ctxs = []
    for r in range(world):
        ctx = M.MsmContext(group, "10")
        ctx.set_bucket_shard(r, world)
        ctx.init_fix_point_list()
        ctx.init_pippenger_CHES_q_over_5()
        ctx.init_pippenger_BGMW95()
        ctxs.append(ctx)
    for method in (1, 2, 3, 4):
        partials = torch.zeros((world, jb), dtype=torch.uint8, device="cuda")
        for r, ctx in enumerate(ctxs):
            ctx.msm_partial_device(method, d_sc.data_ptr(), partials[r].data_ptr())
        torch.cuda.synchronize()
        got = ctxs[0].sum_partials_device(partials.data_ptr(), world)
        assert (got == exp).all(), method
        host = partials.cpu().numpy()
        assert len({host[r].tobytes() for r in range(world)}) > 1  # the work really is split
    for ctx in ctxs:
        ctx.close()


def test_bucket_range_shards_full_size(M, golden):
    import torch

    world = 8
    ctx = M.MsmContext(1, "16")
    ctx.init_fix_point_list()
    ctx.init_pippenger_CHES_q_over_5()
    sc = O.gen_scalars(1, ctx.n)
    d_sc = torch.from_numpy(sc.view(np.uint8).copy()).cuda()
    partials = torch.zeros((world, 144), dtype=torch.uint8, device="cuda")
    for r in range(world):
        ctx.set_bucket_shard(r, world)
        ctx.msm_partial_device(1, d_sc.data_ptr(), partials[r].data_ptr())
        torch.cuda.synchronize()
    ctx.set_bucket_shard(0, 1)
    got = ctx.sum_partials_device(partials.data_ptr(), world)
    assert M.affine_serialize(1, got).hex() == golden["kat_appc"]["g1_n16"]
    ctx.close()
