import sys, os
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, oracle_lib as O, msm_blst_b200 as M
g, cfg = int(sys.argv[1]), sys.argv[2]
ctx = M.MsmContext(g, cfg); ctx.init_fix_point_list(); ctx.init_pippenger_CHES_q_over_5()
sc = O.gen_scalars(1, ctx.n)
for world, rank in ((1, 0), (8, 0), (8, 3), (8, 7), (2, 0), (2, 1)):
    ctx.set_bucket_shard(rank, world)
    for rep in range(3): ctx.msm(1, sc)
    tm = ctx.last_timings()
    print("G%d cfg %s shard %d/%d: dev %.2f ms | %s" % (g, cfg, rank, world, tm["total"], " ".join("%s %.2f" % kv for kv in tm.items() if kv[0] != "total")), flush=True)
