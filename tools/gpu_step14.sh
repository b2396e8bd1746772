#!/bin/bash
# CPM2C head parity on the GPU + the full gpu suite
set -x
O=gpurun_out
timeout 600 python -m pytest tests/test_cpm2c_gpu.py -x -q > $O/r02_cpm2c_tests.log 2>&1; tail -n 15 $O/r02_cpm2c_tests.log
timeout 1500 python -m pytest tests -m gpu -q > $O/r02_gpu_tests_full.log 2>&1; tail -n 8 $O/r02_gpu_tests_full.log
