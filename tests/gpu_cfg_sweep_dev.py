"""Developer probe: same problem (n points) under different reference radices — which (e, h) is best on the GPU for a
shard of that size? (device time of the CHES method; the scalar upload is kept outside the timed events)"""
import sys, os
os.environ["MSMB200_NO_OVERLAP"] = "1"
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle_lib as O
import msm_blst_b200 as M
CASES = ((2, 15, ["13", "15", "16_beta"]), (2, 16, ["15", "16_beta", "16", "17"]), (2, 17, ["15", "16_beta", "16", "17"]), (2, 18, ["16_beta", "16", "18"]),
         (1, 18, ["16", "17", "18"]), (1, 19, ["16", "18", "19"]), (1, 20, ["19", "20", "21"]), (1, 21, ["20_beta", "21"]))
if len(sys.argv) > 1: CASES = [c for c in CASES if str(c[0]) == sys.argv[1]]
for g, nexp, cfgs in CASES:
    n = 1 << nexp
    sc = O.gen_scalars(1, n)
    cf, _ = O.closed_form(g, sc)
    for cfg in cfgs:
        ctx = M.MsmContext(g, cfg, npoints=n)
        ctx.init_fix_point_list(); ctx.init_pippenger_CHES_q_over_5()
        for rep in range(3): r = ctx.msm(1, sc)
        tm = ctx.last_timings()
        print("G%d n=2^%d cfg %-8s e=%d h=%d ok=%s dev %.2f ms | acc %.2f red %.2f fin %.2f" % (g, nexp, cfg, ctx.cfg.e, ctx.cfg.h, (r == cf).all(), tm["total"], tm["accumulate"], tm["reduce"], tm["finalize"]), flush=True)
        ctx.close()
