#!/bin/bash
# One-GPU capture used for the numbers under profiles/ (run on the GPU box via gpurun; every step under its own timeout).
# usage: bash profiles/run_1gpu_capture.sh <tag>      then, here: python profiles/make_ncu_constants.py <tag>
tag=${1:-r2}
out=gpurun_out
M='gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed'
timeout -s KILL 600 python -m pytest tests -m gpu -x -q > $out/${tag}_pytest_gpu.log 2>&1; tail -2 $out/${tag}_pytest_gpu.log
timeout -s KILL 300 python bench.py --steps 20 --warmup 3 > $out/${tag}_bench_N1.json 2> $out/${tag}_bench_N1.err
timeout -s KILL 300 python bench.py --steps 20 --warmup 3 --workload g1_n16 --no-secondary --no-cpu-baseline > $out/${tag}_bench_g1_n16_m1.json 2>> $out/${tag}_bench_N1.err
timeout -s KILL 300 python bench.py --steps 20 --warmup 3 --workload g1_n16 --method 4 --no-secondary --no-cpu-baseline > $out/${tag}_bench_g1_n16_m4.json 2>> $out/${tag}_bench_N1.err
timeout -s KILL 200 msm_blst_b200/csrc/microbench > $out/${tag}_microbench.json 2>&1
timeout -s KILL 200 msm_blst_b200/csrc/microbench gather > $out/${tag}_gather_microbench.json 2>&1
# launch list of the bench command (cold-cache, serialised: compare shares), only after the plain run exited 0
timeout -s KILL 200 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $out/${tag}_plain.log 2>&1 && \
timeout -s KILL 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $out/${tag}_launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $out/${tag}_ncu_launch.log 2>&1
# accumulate-phase kernels of one MSM: duration, DRAM bytes, heavy-FMA-pipe share (-> profiles/r2_ncu_constants.json)
for w in "1 21 g1_n21" "2 18 g2_n18"; do set -- $w
  timeout -s KILL 300 ncu --metrics $M --clock-control none -k regex:'ba_round_kernel|accumulate_kernel' --csv --log-file $out/${tag}_acc_$3.csv \
      python tests/gpu_one_dev.py $1 $2 0 0 3 1 > /dev/null 2>&1
done
# full captures: the four big rounds of the accumulate phase, reduction stage 1 and the finalize (third MSM of the run)
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:'ba_round_kernel' --launch-skip 20 -c 4 -o $out/${tag}_full_ba_g1 \
    python tests/gpu_one_dev.py 1 21 0 0 3 1 > $out/${tag}_ncu_full1.log 2>&1
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:'ba_round_kernel' --launch-skip 14 -c 4 -o $out/${tag}_full_ba_g2 \
    python tests/gpu_one_dev.py 2 18 0 0 3 1 > $out/${tag}_ncu_full2.log 2>&1
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:'list_sum_kernel|list_sum_coop_kernel|bits_finalize_coop|scatter_kernel|ba_emit_kernel|digits_ches' --launch-skip 14 -c 7 -o $out/${tag}_full_rest_g1 \
    python tests/gpu_one_dev.py 1 21 0 0 3 1 > $out/${tag}_ncu_full3.log 2>&1
ls -la $out | grep ${tag}_
