# final 1-GPU validation job (developer helper; every step under its own timeout)
timeout -s KILL 300 python -m pytest tests -m gpu -x -q > gpurun_out/r1e_pytest_gpu.log 2>&1; tail -2 gpurun_out/r1e_pytest_gpu.log
timeout -s KILL 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r1e_smoke.log 2>&1; tail -2 gpurun_out/r1e_smoke.log
timeout -s KILL 300 python bench.py > gpurun_out/r1e_bench_g1_n21_N1.json 2> gpurun_out/r1e_bench_g1_n21_N1.err
timeout -s KILL 300 python bench.py --workload g2_n18 > gpurun_out/r1e_bench_g2_n18_N1.json 2> gpurun_out/r1e_bench_g2_n18_N1.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r1e_bench_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, round(d["value"],3), round(d["e2e"]["value"],3), {k:round(v,2) for k,v in d.get("phases_ms",{}).items()}, d["gpu_launches"], d["roofline"]["frac"], d["cpu_baseline"]["value"])
    except Exception as e: print(f, "ERR", e)
PY
