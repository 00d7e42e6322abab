for args in "2 18 0 0 3 1" "2 15 0 0 3 1" "2 16 0 0 3 1" "2 13 0 0 3 1" "2 18 1 0 3 1" "2 20 0 0 3 1"; do timeout -s KILL 120 python tests/gpu_one_dev.py $args 2>&1 | tail -1; done
timeout -s KILL 600 python -m pytest tests -m gpu -x -q -k "g2 or G2 or fp2 or [2]" 2>&1 | tail -3
