"""Developer probe: ONE configuration, a few repetitions (the command line ncu wraps).
usage: gpu_one_dev.py <group> <config> <accum 1|2> <ba_batch_max> <reps> [method] [tuning_key=value ...]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle_lib as O
import msm_blst_b200 as M
import msm_blst_b200.api as _A
if os.environ.get("MSMB200_DEV_LIB"): _A.LIB_PATH = os.path.abspath(os.environ["MSMB200_DEV_LIB"])
g, cfg, accum, bmax, reps = int(sys.argv[1]), sys.argv[2], int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5])
method = int(sys.argv[6]) if len(sys.argv) > 6 and "=" not in sys.argv[6] else 1
tune = [a.split("=") for a in sys.argv[6:] if "=" in a]
ctx = M.MsmContext(g, cfg); ctx.init_fix_point_list()
if method in (1, 2): ctx.init_pippenger_CHES_q_over_5()
if method == 3: ctx.init_pippenger_BGMW95()
sc = O.gen_scalars(1, ctx.n); cf, _ = O.closed_form(g, sc)
ctx.set_accumulator(accum)
if bmax: ctx.set_tuning("ba_batch_max", bmax)
for k, v in tune: ctx.set_tuning(k, int(v))
for rep in range(reps): r = ctx.msm(method, sc)
tm = ctx.last_timings()
tstr = " ".join("%s=%s" % (k, v) for k, v in tune)
print("G%d cfg %s m%d accum %d bmax %2d %s ok=%s total %.2f | dig %.2f sort %.2f acc %.2f red %.2f fin %.2f launches %d" % (g, cfg, method, accum, bmax, tstr, (r == cf).all(), tm["total"], tm["digits"], tm["sort"], tm["accumulate"], tm["reduce"], tm["finalize"], ctx.last_launches()), flush=True)
ctx.close()
