#!/bin/bash
set -x
O=gpurun_out
( timeout 300 python tools/rn50_throughput.py 8 6; SPM_RN50_FRONT_CHUNK=108 timeout 300 python tools/rn50_throughput.py 8 6 ) 2>&1 | grep -v "^+" > $O/r02_rn50_s7.log; cat $O/r02_rn50_s7.log
R="python tools/profile_rn50.py 216 3"
timeout 300 $R > $O/r02_s7_plain.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"gemm_tcgen05_kernel|gemm2_tcgen05" -s 230 -c 13 -o $O/r02_rn50_front $R > $O/r02_s7_ncu.log 2>&1
ncu -i $O/r02_rn50_front.ncu-rep --page raw --csv > $O/r02_rn50_front_raw.csv 2>/dev/null
ncu -i $O/r02_rn50_front.ncu-rep --page source --csv > $O/r02_rn50_front_source.csv 2>/dev/null
rm -f $O/r02_rn50_front.ncu-rep
cat $O/r02_s7_plain.log
du -sh $O
