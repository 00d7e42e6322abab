"""Developer probe: same problem (n points) under different reference radices — which (e, h) is best on the GPU?"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import oracle_lib as O
import msm_blst_b200 as M
for g, nexp, cfgs in ((1, 21, ["21", "20_beta", "17_beta", "16_beta", "13"]), (2, 18, ["18", "17_beta", "16_beta", "13"]), (1, 18, ["18", "16_beta", "13"])):
    n = 1 << nexp
    sc = O.gen_scalars(1, n)
    cf, _ = O.closed_form(g, sc)
    for cfg in cfgs:
        ctx = M.MsmContext(g, cfg, npoints=n)
        ctx.init_fix_point_list(); ctx.init_pippenger_CHES_q_over_5()
        for rep in range(3): r = ctx.msm(1, sc)
        tm = ctx.last_timings()
        print("G%d n=2^%d cfg %-8s e=%d h=%d ok=%s dev %.2f ms | acc %.2f red %.2f" % (g, nexp, cfg, ctx.cfg.e, ctx.cfg.h, (r == cf).all(), tm["total"], tm["accumulate"], tm["reduce"]), flush=True)
        ctx.close()
