"""Developer soak: many batch-affine MSMs on fresh scalar sets, each compared with the closed form (rare-race hunting).
usage: gpu_soak_dev.py <seconds> [group:config ...]"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import oracle_lib as O
import msm_blst_b200 as M
budget = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
cases = [(1, "16"), (2, "13"), (1, "18"), (2, "15"), (1, "13"), (2, "16")]
if len(sys.argv) > 2: cases = [(int(a.split(":")[0]), a.split(":")[1]) for a in sys.argv[2:]]
ctxs = []
for g, cfg in cases:
    c = M.MsmContext(g, cfg); c.init_fix_point_list(); c.init_pippenger_CHES_q_over_5(); c.init_pippenger_BGMW95(); c.set_accumulator(2)
    ctxs.append(c)
t0 = time.time(); runs = 0; bad = 0; seed = 5000
while time.time() - t0 < budget:
    for (g, cfg), c in zip(cases, ctxs):
        seed += 1
        sc = O.gen_scalars(seed, c.n)
        if seed % 7 == 0: sc[: c.n // 2] = sc[0]          # half of the scalars equal: P + P slots, heavy buckets
        if seed % 11 == 0: sc[1::2] = 0                     # every other scalar zero
        exp, _ = O.closed_form(g, sc)
        for method in (1, 3):
            got = c.msm(method, sc)
            runs += 1
            if not (got == exp).all():
                bad += 1
                print("MISMATCH group %d cfg %s method %d seed %d" % (g, cfg, method, seed), flush=True)
print("soak: %d MSMs in %.0f s, %d mismatches" % (runs, time.time() - t0, bad))
sys.exit(1 if bad else 0)
