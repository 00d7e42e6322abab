timeout -s KILL 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for args in "1 21 0 0 3 1" "2 18 0 0 3 1" "1 16 0 0 3 1" "1 18 0 0 3 1" "2 15 0 0 3 1" "1 16 0 0 3 4"; do timeout -s KILL 120 python tests/gpu_one_dev.py $args 2>&1 | tail -1; done
timeout -s KILL 100 python tests/gpu_perf_dev.py 1:10 2>&1 | grep "fixpts\|method 1"
