for args in "1 21 2 0 3 1" "1 21 2 0 3 1 ba_batch_max=160" "1 21 2 0 3 1 ba_batch_max=80" "2 18 2 0 3 1" "1 16 2 0 3 1" "1 19 2 0 3 1" "1 20 2 0 3 1" "2 16 2 0 3 1" "2 20 2 0 3 1"; do timeout -s KILL 120 python tests/gpu_one_dev.py $args 2>&1 | tail -1; done
timeout -s KILL 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
