#!/bin/bash
O=gpurun_out
echo "== tests rn50"; timeout 900 python -m pytest tests -m gpu -x -q -k "rn50" 2>&1 | tail -n 4
( timeout 300 python tools/rn50_throughput.py 8 6; SPM_CONV_WIN=0 timeout 300 python tools/rn50_throughput.py 8 6 ) 2>&1 | grep -v "^+" | cut -c40-
R="python tools/profile_rn50.py 320 4"
timeout 300 $R > $O/r02_s13_plain.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 170 -c 85 --csv --log-file $O/r02_launches_rn50_s13.csv $R > $O/r02_s13_ncu.log 2>&1
cat $O/r02_s13_plain.log
echo "== jpeg sweep test"; timeout 600 python -m pytest tests -m gpu -x -q -k "jpeg" 2>&1 | tail -n 4
