#!/bin/bash
# ncu --set full with source counters of the current fused OTAM kernel (P = 1000)
set -x
O=gpurun_out
T="python tools/time_head_kernels.py --one"
SPM_OTAM_FUSED=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"otam_fused" -s 3 -c 1 -o $O/r02_s22_otamf $T > $O/r02_s22_ncu.log 2>&1
ncu -i $O/r02_s22_otamf.ncu-rep --page raw --csv > $O/r02_s22_otamf_raw.csv 2>/dev/null
ncu -i $O/r02_s22_otamf.ncu-rep --page source --csv > $O/r02_s22_otamf_source.csv 2>/dev/null
rm -f $O/r02_s22_otamf.ncu-rep
tail -n 3 $O/r02_s22_ncu.log
