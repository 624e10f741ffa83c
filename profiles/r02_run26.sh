# pass 26 (1 GPU): DBSCAN labels (range-dependent radius searches on the grid + the reference's growth loop); whole suite
timeout 600 python -m pytest tests/test_prep_stages.py -m gpu -x -q 2>&1 | tail -12
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
