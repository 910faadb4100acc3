timeout 200 python tests/debug_tc_stats.py 2>&1 | tail -4
timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "layer" 2>&1 | tail -2
