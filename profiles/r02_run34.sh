# pass 34 (1 GPU): pool tests with two registrations per launch; the whole suite
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "pool or batch" 2>&1 | tail -30
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
