#!/bin/bash
# full GPU suite on the current build (CLIP-FSAR optional branches, LayerNorm folding) + head-kernel timings at batch scale
set -x
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > $O/r02_s17_tests.log 2>&1; tail -n 15 $O/r02_s17_tests.log
timeout 300 python tools/time_head_kernels.py > $O/r02_s17_head_kernels.log 2>&1; tail -n 30 $O/r02_s17_head_kernels.log
