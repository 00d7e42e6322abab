"""Developer probe (library built with -DMSMB200_BA_TIMING): per-round phase durations of the batch-affine kernel."""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import oracle_lib as O
import msm_blst_b200 as M
import msm_blst_b200.api as _A
if os.environ.get("MSMB200_DEV_LIB"): _A.LIB_PATH = os.path.abspath(os.environ["MSMB200_DEV_LIB"])
g, cfg, bmax = int(sys.argv[1]), sys.argv[2], int(sys.argv[3])
ctx = M.MsmContext(g, cfg); ctx.init_fix_point_list(); ctx.init_pippenger_CHES_q_over_5()
sc = O.gen_scalars(1, ctx.n)
ctx.set_accumulator(2); ctx.set_tuning("ba_batch_max", bmax)
for rep in range(3): r = ctx.msm(1, sc)
print(ctx.last_timings())
dbg = np.zeros((32, 4096, 4), dtype=np.uint64)
assert M.lib().msmb200_debug_ba_timing(C.c_void_p(dbg.ctypes.data)) == 0
for r in range(12):
    d = dbg[r]
    live = d[:, 3] > 0
    if not live.any(): continue
    d = d[live].astype(np.int64)
    fwd, inv, bwd = d[:, 1] - d[:, 0], d[:, 2] - d[:, 1], d[:, 3] - d[:, 2]
    print("round %2d blocks(<=4096) %5d  fwd %9.0f  inv %9.0f  bwd %9.0f cycles (mean per block); total %9.0f; spread of block totals min %9.0f max %9.0f" % (
        r, live.sum(), fwd.mean(), inv.mean(), bwd.mean(), (d[:, 3] - d[:, 0]).mean(), (d[:, 3] - d[:, 0]).min(), (d[:, 3] - d[:, 0]).max()))
