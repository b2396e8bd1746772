#!/bin/bash
set -x
O=gpurun_out
SPM_OTAM_TC_MINP=296 timeout 300 python tools/otam_dp_check.py > $O/r02_s23_check.log 2>&1; grep "P=300 W=5 Q=5 T=8\|P=1974\|P=1800\|worst" $O/r02_s23_check.log | cut -c1-150
SPM_OTAM_TC_MINP=296 timeout 300 python tools/time_head_kernels.py 2>&1 | tail -n 4 > $O/r02_s23_head_kernels.log 2>&1
cat $O/r02_s23_head_kernels.log
timeout 300 python tools/time_jpeg.py > $O/r02_s23_jpeg.log 2>&1; tail -n 2 $O/r02_s23_jpeg.log
