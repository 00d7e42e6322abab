for args in "1 21 2 0 3 1" "1 21 1 0 3 1" "2 18 2 0 3 1" "1 16 0 0 3 1" "1 19 0 0 3 1" "1 20 2 0 3 1" "1 18 1 0 3 1" "2 15 1 0 3 1" "1 21 2 0 3 3"; do timeout -s KILL 120 python tests/gpu_one_dev.py $args 2>&1 | tail -1; done
MSMB200_PACKED_TABLES=1 timeout -s KILL 120 python tests/gpu_one_dev.py 1 21 2 0 3 1 2>&1 | tail -1
timeout -s KILL 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed --clock-control none -k regex:ba_round_kernel --csv --log-file gpurun_out/r2ac_rounds.csv python tests/gpu_one_dev.py 1 21 2 0 1 1 > /dev/null 2>&1
