"""Developer smoke script (run on the GPU box): field/point parity + C10 G1/G2 all methods vs oracle."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import oracle_lib as O
import msm_blst_b200 as M

o = O.oracle()
rng = np.random.default_rng(7)
def rand_fp(n):
    a = rng.integers(0, 2**64, size=(n, 6), dtype=np.uint64)
    a[:, 5] &= np.uint64(0x0fffffffffffffff)
    return a
n = 4096
a, b = rand_fp(n), rand_fp(n)
for op in range(6):
    exp = np.empty_like(a); o.oracle_fp_op(op, O.ptr(a), O.ptr(b), O.ptr(exp), n)
    got = M.test_field_op(1, op, a, b)
    print("fp op", op, (got == exp).all())
exp = np.empty_like(a[:64]); o.oracle_fp_op(6, O.ptr(a[:64].copy()), O.ptr(b), O.ptr(exp), 64)
print("fp inv", (M.test_field_op(1, 6, a[:64].copy()) == exp).all())
a2, b2 = rand_fp(2 * n).reshape(n, 12), rand_fp(2 * n).reshape(n, 12)
for op in range(6):
    exp = np.empty_like(a2); o.oracle_fp2_op(op, O.ptr(a2), O.ptr(b2), O.ptr(exp), n)
    got = M.test_field_op(2, op, a2, b2)
    print("fp2 op", op, (got == exp).all())

sc = O.gen_scalars(1, 1024)
for g in (1, 2):
    t = time.time()
    ctx = M.MsmContext(g, "10")
    ctx.init_fix_point_list()
    oc = O.OracleCtx(g, "10", threads=8); oc.init_fix_points()
    print("G%d fix points equal:" % g, (ctx.download(0) == oc.points()).all(), time.time() - t)
    t = time.time(); ctx.init_pippenger_CHES_q_over_5(); ctx.init_pippenger_BGMW95(); print("gpu tables", time.time() - t)
    t = time.time(); oc.build_table(0); oc.build_table(1); print("cpu tables", time.time() - t)
    print("table 3nh equal:", (ctx.download(1) == oc.table(0)).all(), " bgmw equal:", (ctx.download(2) == oc.table(1)).all())
    cf, k = O.closed_form(g, sc)
    for m in (1, 2, 3, 4):
        r = ctx.msm(m, sc)
        e = oc.msm(m, sc)
        print("G%d method %d: gpu==oracle %s gpu==closed %s" % (g, m, (r == e).all(), (r == cf).all()), ctx.last_timings(), ctx.last_launches())
    print("serialize", M.affine_serialize(g, r).hex()[:32], O.serialize(g, cf).hex()[:32])
