# 1-GPU validation job (developer helper; every step under its own timeout)
timeout -s KILL 300 python -m pytest tests -m gpu -x -q > gpurun_out/r1d_pytest_gpu.log 2>&1; tail -2 gpurun_out/r1d_pytest_gpu.log
timeout -s KILL 300 python bench.py --steps 20 --warmup 3 > gpurun_out/r1d_bench_g1_n21_N1.json 2> gpurun_out/r1d_bench_g1_n21_N1.err
timeout -s KILL 300 python bench.py --steps 20 --warmup 3 --workload g2_n18 > gpurun_out/r1d_bench_g2_n18_N1.json 2> gpurun_out/r1d_bench_g2_n18_N1.err
timeout -s KILL 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --workload g1_n16 > gpurun_out/r1d_bench_g1_n16_m1.json 2>/dev/null
for m in 2 3 4; do timeout -s KILL 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --workload g1_n16 --method $m > gpurun_out/r1d_bench_g1_n16_m$m.json 2>/dev/null; done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r1d_bench_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, round(d["value"],3), round(d["e2e"]["value"],3), {k:round(v,2) for k,v in d.get("phases_ms",{}).items()})
    except Exception as e: print(f, "ERR", e)
PY
