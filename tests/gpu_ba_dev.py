"""Developer probe: XYZZ work items vs batch-affine rounds at G1 n=2^21 / G2 n=2^18, batch-size / stagger sweep."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle_lib as O
import msm_blst_b200 as M
import msm_blst_b200.api as _A
if os.environ.get("MSMB200_DEV_LIB"): _A.LIB_PATH = os.path.abspath(os.environ["MSMB200_DEV_LIB"])  # developer builds of the same library
print("lib", _A.LIB_PATH, flush=True)
which = sys.argv[1:] or ["1:21", "2:18"]
for spec in which:
    g, cfg = spec.split(":")
    g = int(g)
    ctx = M.MsmContext(g, cfg); ctx.init_fix_point_list(); ctx.init_pippenger_CHES_q_over_5()
    sc = O.gen_scalars(1, ctx.n); cf, _ = O.closed_form(g, sc)
    for accum, bmax, stag in ((1, 0, 0), (2, 64, 1), (2, 110, 1)):
        ctx.set_accumulator(accum)
        if bmax: ctx.set_tuning("ba_batch_max", bmax); ctx.set_tuning("ba_stagger", stag)
        for rep in range(3): r = ctx.msm(1, sc)
        tm = ctx.last_timings()
        print("G%d cfg %s accum %d bmax %3d stagger %d ok=%s total %.2f | dig %.2f sort %.2f acc %.2f red %.2f fin %.2f launches %d" % (g, cfg, accum, bmax, stag, (r == cf).all(), tm["total"], tm["digits"], tm["sort"], tm["accumulate"], tm["reduce"], tm["finalize"], ctx.last_launches()), flush=True)
    ctx.close()
