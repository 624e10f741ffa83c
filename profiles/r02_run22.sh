# pass 22 (1 GPU): FastVGICP (vgicp.cu) against the oracle; the whole GPU suite
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_real_data.py tests/test_shim.py -m gpu -x -q -k "vgicp or shim" 2>&1 | tail -25
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
