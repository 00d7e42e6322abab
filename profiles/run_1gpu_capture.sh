#!/bin/bash
# One-GPU capture used for the numbers under profiles/ (run on the GPU box via gpurun; every step under its own timeout).
# usage: bash profiles/run_1gpu_capture.sh <tag>
tag=${1:-r1c}
out=gpurun_out
timeout -s KILL 300 python -m pytest tests -m gpu -x -q > $out/${tag}_pytest_gpu.log 2>&1; tail -2 $out/${tag}_pytest_gpu.log
timeout -s KILL 300 python bench.py --steps 20 --warmup 3 > $out/${tag}_bench_g1_n21_N1.json 2> $out/${tag}_bench_g1_n21_N1.err
timeout -s KILL 300 python bench.py --steps 20 --warmup 3 --workload g2_n18 > $out/${tag}_bench_g2_n18_N1.json 2> $out/${tag}_bench_g2_n18_N1.err
# launch list of the bench command (cold-cache, serialised: compare shares)
timeout -s KILL 200 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $out/${tag}_plain.log 2>&1 && \
timeout -s KILL 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $out/${tag}_launches_g1_n21.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $out/${tag}_ncu_launch.log 2>&1
timeout -s KILL 200 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --workload g2_n18 > $out/${tag}_plain2.log 2>&1 && \
timeout -s KILL 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $out/${tag}_launches_g2_n18.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --workload g2_n18 > $out/${tag}_ncu_launch2.log 2>&1
# full capture of the dominant kernel and of the reduction kernels (one launch each)
timeout -s KILL 400 ncu --set full --clock-control none --import-source on -k regex:'accumulate_kernel|list_sum_kernel|list_sum_coop_kernel|bits_finalize_coop' \
    --launch-skip 12 -c 6 -o $out/${tag}_full_g1 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $out/${tag}_ncu_full.log 2>&1
ls -la $out | grep $tag
