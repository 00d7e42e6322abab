"""Developer probe for an ncu launch list: a few MSMs of one (group, config, method) after setup."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle_lib as O
import msm_blst_b200 as M

g, cfg = sys.argv[1].split(":")
method = int(sys.argv[2]) if len(sys.argv) > 2 else 1
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
ctx = M.MsmContext(int(g), cfg)
ctx.init_fix_point_list()
if method in (1, 2):
    ctx.init_pippenger_CHES_q_over_5()
if method == 3:
    ctx.init_pippenger_CHES_q_over_5()
    ctx.init_pippenger_BGMW95()
sc = O.gen_scalars(1, ctx.n)
for rep in range(reps):
    r = ctx.msm(method, sc)
print(sys.argv[1], "method", method, ctx.last_timings(), "launches", ctx.last_launches())
ctx.close()
