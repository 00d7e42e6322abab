for args in "2 18 2 0 3 1" "2 18 1 0 3 1" "2 16 2 0 3 1" "2 20 2 0 3 1" "2 21 2 0 3 1" "1 21 2 0 3 1"; do timeout -s KILL 120 python tests/gpu_one_dev.py $args 2>&1 | tail -1; done
timeout -s KILL 600 python -m pytest tests -m gpu -x -q -k "g2 or G2 or accumulator or reducer or [2]" 2>&1 | tail -3
timeout -s KILL 600 python tests/gpu_cfg_sweep_dev.py 2>&1 | grep "^G"
