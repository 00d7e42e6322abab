"""Developer probe: XYZZ work items (accumulator 1) vs batch-affine rounds (2) over sizes and methods."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle_lib as O
import msm_blst_b200 as M
for g, cfg, methods in ((1, "14", (1,)), (1, "16", (1, 3, 4)), (1, "18", (1,)), (1, "19", (1,)), (1, "20", (1,)), (1, "21", (1, 3, 4)), (2, "16", (1,)), (2, "18", (1, 3, 4)), (2, "20", (1,))):
    ctx = M.MsmContext(g, cfg); ctx.init_fix_point_list()
    if 1 in methods: ctx.init_pippenger_CHES_q_over_5()
    if 3 in methods: ctx.init_pippenger_BGMW95()
    sc = O.gen_scalars(1, ctx.n); cf, _ = O.closed_form(g, sc)
    for method in methods:
        for accum in (1, 2):
            ctx.set_accumulator(accum)
            for rep in range(3): r = ctx.msm(method, sc)
            tm = ctx.last_timings()
            print("G%d cfg %-3s m%d accum %d ok=%s total %.3f | dig %.3f sort %.3f acc %.3f red %.3f fin %.3f launches %d" % (g, cfg, method, accum, (r == cf).all(), tm["total"], tm["digits"], tm["sort"], tm["accumulate"], tm["reduce"], tm["finalize"], ctx.last_launches()), flush=True)
    ctx.close()
