timeout 300 python -m pytest tests/test_gemm_gpu.py -x -q 2>&1 | tail -2
timeout 300 python tools/time_small_gemm.py 2>&1 | tail -13
timeout 300 python tools/rn50_throughput.py 2>&1 | tail -1
