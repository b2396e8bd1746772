#!/bin/bash
set -x
O=gpurun_out
timeout 600 python -m pytest tests/test_optim_gpu.py -q > $O/r02_s29_optim.log 2>&1; tail -n 3 $O/r02_s29_optim.log
timeout 900 python -m pytest tests/test_stages_gpu.py -q -k "otam" > $O/r02_s29_otam.log 2>&1; tail -n 3 $O/r02_s29_otam.log
timeout 300 python tools/time_head_kernels.py 2>&1 | tail -n 4
