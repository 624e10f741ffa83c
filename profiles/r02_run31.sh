# pass 31 (1 GPU): the CUDA library against the fixtures written by the reference's own code (registration + DBSCAN)
timeout 900 python -m pytest tests/test_reference_code.py tests/test_prep_stages.py -m gpu -x -q 2>&1 | tail -15
