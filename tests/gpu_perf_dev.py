"""Developer perf probe (GPU box): per-phase timings at the BASELINE sizes, result checked against the closed form."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import oracle_lib as O
import msm_blst_b200 as M

cases = sys.argv[1:] or ["1:16", "1:21", "2:18"]
for case in cases:
    g, cfg = case.split(":")
    g = int(g)
    t = time.time(); ctx = M.MsmContext(g, cfg); ctx.init_fix_point_list(); t_fix = time.time() - t
    n = ctx.n
    t = time.time(); ctx.init_pippenger_CHES_q_over_5(); t_ches = time.time() - t
    t = time.time(); ctx.init_pippenger_BGMW95(); t_bg = time.time() - t
    print("G%d cfg %s n=%d fixpts %.2fs table3nh %.2fs bgmw %.2fs" % (g, cfg, n, t_fix, t_ches, t_bg), flush=True)
    sc = O.gen_scalars(1, n)
    t = time.time(); cf, k = O.closed_form(g, sc); print("closed form %.2fs k=%x" % (time.time() - t, k))
    for m in (1, 2, 3, 4):
        for rep in range(3):
            t = time.time(); r = ctx.msm(m, sc); wall = time.time() - t
        tm = ctx.last_timings()
        print("  method %d ok=%s wall %.2f ms dev %.2f ms | %s | launches %d" % (
            m, (r == cf).all(), wall * 1e3, tm["total"], " ".join("%s %.2f" % (k2, v) for k2, v in tm.items() if k2 != "total"), ctx.last_launches()), flush=True)
    print("  result", M.affine_serialize(g, r).hex())
    ctx.close()
