#!/bin/bash
# parity + timing of the r02 kernels (fused persistent OTAM, warp-wavefront soft-DTW)
set -x
O=gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "otam or softdtw" > $O/r02_newk_tests.log 2>&1; tail -n 5 $O/r02_newk_tests.log
timeout 300 python tools/time_head_kernels.py > $O/r02_otam_fused_times.log 2>&1; cat $O/r02_otam_fused_times.log
SPM_OTAM_FUSED=0 timeout 300 python tools/time_head_kernels.py > $O/r02_otam_twokernel_times.log 2>&1; cat $O/r02_otam_twokernel_times.log
timeout 600 python tools/time_softdtw.py > $O/r02_softdtw_times.log 2>&1; cat $O/r02_softdtw_times.log
