#!/bin/bash
set -x
O=gpurun_out
export_rep() { ncu -i $O/$1.ncu-rep --page raw --csv > $O/$1_raw.csv 2>/dev/null; [ "$2" = "source" ] && ncu -i $O/$1.ncu-rep --page source --csv > $O/$1_source.csv 2>/dev/null; rm -f $O/$1.ncu-rep; }
LIGHT="--section SpeedOfLight --section MemoryWorkloadAnalysis --section LaunchStats --section Occupancy --section WarpStateStats --section SchedulerStats"
timeout 600 python -m pytest tests -m gpu -x -q -k "otam or softdtw" > $O/r02_newk_tests.log 2>&1; tail -n 5 $O/r02_newk_tests.log
timeout 120 python tools/repro_softdtw.py 4096 40 38
timeout 300 python tools/time_head_kernels.py > $O/r02_otam_fused_times.log 2>&1; cat $O/r02_otam_fused_times.log
timeout 600 python tools/time_softdtw.py > $O/r02_softdtw_times.log 2>&1; cat $O/r02_softdtw_times.log
T="python tools/time_head_kernels.py --one"
timeout 300 $T > $O/r02_prof_otamf_plain.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"otam_fused" -s 3 -c 2 -o $O/r02_otam_fused $T > $O/r02_prof_otamf_ncu.log 2>&1
export_rep r02_otam_fused source
R="python tools/rn50_throughput.py"
timeout 300 $R > $O/r02_prof_rn50_plain.log 2>&1 &&
timeout 900 ncu $LIGHT --clock-control none -k regex:"gemm_tcgen05_kernel|gemm2_tcgen05" -s 660 -c 66 -o $O/r02_rn50_convs $R > $O/r02_prof_rn50c_ncu.log 2>&1
export_rep r02_rn50_convs
du -sh $O
