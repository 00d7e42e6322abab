"""Developer probe: batch-affine accumulator variants at G1 n=2^21 / G2 n=2^18 (env knobs read per call)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle_lib as O
import msm_blst_b200 as M
for g, cfg in ((1, "21"), (2, "18")):
    ctx = M.MsmContext(g, cfg); ctx.init_fix_point_list(); ctx.init_pippenger_CHES_q_over_5()
    sc = O.gen_scalars(1, ctx.n); cf, _ = O.closed_form(g, sc)
    for accum, batch in ((1, 0), (2, 4), (2, 8), (2, 16), (2, 32)):
        os.environ["MSMB200_ACCUM"] = str(accum)
        if batch: os.environ["MSMB200_BA_BATCH"] = str(batch)
        for rep in range(3): r = ctx.msm(1, sc)
        tm = ctx.last_timings()
        print("G%d cfg %s accum %d batch %2d ok=%s total %.2f | acc %.2f red %.2f fin %.2f" % (g, cfg, accum, batch, (r == cf).all(), tm["total"], tm["accumulate"], tm["reduce"], tm["finalize"]), flush=True)
    ctx.close()
