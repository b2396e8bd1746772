#!/bin/bash
set -x
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "rn50 or otam or gemm" > $O/r02_s5_tests.log 2>&1; tail -n 6 $O/r02_s5_tests.log
timeout 300 python tools/time_head_kernels.py > $O/r02_otam_fused_times2.log 2>&1; cat $O/r02_otam_fused_times2.log
R="python tools/rn50_throughput.py 8 6"
( timeout 300 $R
SPM_CONV_2CTA=0 timeout 300 $R
SPM_CONV_NARROW=0 timeout 300 $R
SPM_CONV_2CTA=0 SPM_CONV_NARROW=0 SPM_RN50_FRONT_CHUNK=64 SPM_RN50_BACK_CHUNK=64 timeout 300 $R
SPM_RN50_FRONT_CHUNK=36 SPM_RN50_BACK_CHUNK=216 timeout 300 $R
SPM_RN50_FRONT_CHUNK=108 SPM_RN50_BACK_CHUNK=216 timeout 300 $R
SPM_RN50_FRONT_CHUNK=72 SPM_RN50_BACK_CHUNK=432 timeout 300 $R
SPM_RN50_FRONT_CHUNK=72 SPM_RN50_BACK_CHUNK=144 timeout 300 $R ) 2>&1 | grep -v "^+" > $O/r02_rn50_ab.log
cat $O/r02_rn50_ab.log
timeout 300 $R > $O/r02_prof_rn50_plain.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 1300 -c 500 --csv --log-file $O/r02_launches_rn50_new.csv $R > $O/r02_prof_rn50_ncu.log 2>&1
du -sh $O
