#!/bin/bash
set -x
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "rn50 or gemm" > $O/r02_s8_tests.log 2>&1; tail -n 4 $O/r02_s8_tests.log
( timeout 300 python tools/rn50_throughput.py 8 6; SPM_RN50_FRONT_CHUNK=72 timeout 300 python tools/rn50_throughput.py 8 6; SPM_RN50_FRONT_CHUNK=216 timeout 300 python tools/rn50_throughput.py 8 6; SPM_RN50_BACK_CHUNK=432 timeout 300 python tools/rn50_throughput.py 8 6 ) 2>&1 | grep -v "^+" > $O/r02_rn50_s8.log; cat $O/r02_rn50_s8.log
R="python tools/profile_rn50.py 216 3"
timeout 300 $R > $O/r02_s8_plain.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 160 --csv --log-file $O/r02_launches_rn50_s8.csv $R > $O/r02_s8_ncu.log 2>&1
cat $O/r02_s8_plain.log
