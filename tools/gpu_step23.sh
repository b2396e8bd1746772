#!/bin/bash
set -x
O=gpurun_out
timeout 900 python -m pytest tests/test_stages_gpu.py tests/test_canaries_gpu.py -q -k "otam or canar" > $O/r02_s23_tests.log 2>&1; tail -n 8 $O/r02_s23_tests.log
timeout 300 python tools/time_head_kernels.py 2>&1 | tail -n 4 > $O/r02_s23_head_kernels.log 2>&1
cat $O/r02_s23_head_kernels.log
