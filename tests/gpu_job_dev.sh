timeout -s KILL 120 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "wbits or points_add or pippenger" > gpurun_out/tests_wbits.log 2>&1; tail -5 gpurun_out/tests_wbits.log
timeout -s KILL 100 python tests/gpu_perf_dev.py 1:21 > gpurun_out/perf_l2fetch.log 2>&1; grep "method 1" gpurun_out/perf_l2fetch.log
MSMB200_L2_FETCH=0 timeout -s KILL 100 python tests/gpu_perf_dev.py 1:21 > gpurun_out/perf_l2fetch0.log 2>&1; grep "method 1" gpurun_out/perf_l2fetch0.log
timeout -s KILL 100 python tests/gpu_stage_dev.py 1:21 1 2 > gpurun_out/l2f_plain.log 2>&1 && timeout -s KILL 200 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:accumulate_kernel -c 2 --csv --log-file gpurun_out/l2f_acc.csv python tests/gpu_stage_dev.py 1:21 1 2 > gpurun_out/l2f_ncu.log 2>&1
grep -v "^==" gpurun_out/l2f_acc.csv | cut -d, -f5,13- | tail -9
