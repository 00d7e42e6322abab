# fast-squaring validation job (developer helper; every step under its own timeout)
timeout -s KILL 60 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "fp_ops or point_ops or accumulator_modes or reducer_modes" > gpurun_out/sqr_t1.log 2>&1; tail -2 gpurun_out/sqr_t1.log
timeout -s KILL 70 python -m pytest tests/test_gpu_full_sizes.py -m gpu -x -q -k "known_answer or other_configs or all_reference_configs" > gpurun_out/sqr_t2.log 2>&1; tail -2 gpurun_out/sqr_t2.log
timeout -s KILL 40 python tests/gpu_perf_dev.py 1:21 > gpurun_out/sqr_perf.log 2>&1; grep "method" gpurun_out/sqr_perf.log
