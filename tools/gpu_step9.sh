#!/bin/bash
set -x
O=gpurun_out
( timeout 300 python tools/rn50_throughput.py 8 6; SPM_RN50_FRONT_CHUNK=432 SPM_RN50_BACK_CHUNK=432 timeout 300 python tools/rn50_throughput.py 8 6; SPM_RN50_FRONT_CHUNK=320 SPM_RN50_BACK_CHUNK=320 timeout 300 python tools/rn50_throughput.py 8 6 ) 2>&1 | grep -v "^+" > $O/r02_rn50_s9.log; cat $O/r02_rn50_s9.log
R="python tools/profile_rn50.py 216 4"
timeout 300 $R > $O/r02_s9_plain.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 160 -c 80 --csv --log-file $O/r02_launches_rn50_s9.csv $R > $O/r02_s9_ncu.log 2>&1
cat $O/r02_s9_plain.log
