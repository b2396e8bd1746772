#!/bin/bash
# fused OTAM kernel: L2 bulk prefetch 0/1/2 problems ahead; ncu --set full with source counters of the fused kernel
set -x
O=gpurun_out
for v in "SPM_OTAM_PF=0" "SPM_OTAM_PF=1" "SPM_OTAM_PF=2" "SPM_OTAM_PF=3"; do
  echo "== [$v]"; env SPM_OTAM_FUSED=1 $v timeout 300 python tools/time_head_kernels.py 2>&1 | tail -n 4
done > $O/r02_s19_head_kernels.log 2>&1
cat $O/r02_s19_head_kernels.log
T="python tools/time_head_kernels.py --one"
SPM_OTAM_FUSED=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"otam_fused" -s 3 -c 1 -o $O/r02_s19_otamf $T > $O/r02_s19_ncu.log 2>&1
ncu -i $O/r02_s19_otamf.ncu-rep --page raw --csv > $O/r02_s19_otamf_raw.csv 2>/dev/null
ncu -i $O/r02_s19_otamf.ncu-rep --page source --csv > $O/r02_s19_otamf_source.csv 2>/dev/null
rm -f $O/r02_s19_otamf.ncu-rep
tail -n 3 $O/r02_s19_ncu.log
