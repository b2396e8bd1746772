#!/bin/bash
# ncu evidence of the r02 build (run under gpurun; every ncu run follows a plain run of the same command that exited 0).
# The .ncu-rep files stay on the box: gpurun brings back at most 64 MiB, so each report is exported to CSV pages here.
set -x
O=gpurun_out
export_rep() {  # $1 = report stem, $2 = "source" to export the source page as well
  ncu -i $O/$1.ncu-rep --page raw --csv > $O/$1_raw.csv 2>/dev/null
  [ "$2" = "source" ] && ncu -i $O/$1.ncu-rep --page source --csv > $O/$1_source.csv 2>/dev/null
  rm -f $O/$1.ncu-rep
}
LIGHT="--section SpeedOfLight --section MemoryWorkloadAnalysis --section LaunchStats --section Occupancy --section WarpStateStats"
B="python bench.py --steps 2 --warmup 3 --no-e2e --no-extra-legs --no-cpu-baseline"
timeout 300 $B > $O/r02_prof_bench_plain.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 2500 -c 1700 --csv --log-file $O/r02_launches_vit_16eps.csv $B > $O/r02_prof_bench_ncu.log 2>&1
E="python tools/profile_encoder.py 512 3"
timeout 300 $E > $O/r02_prof_enc_plain.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"gemm2_tcgen05|vit_attention_tc|layernorm" -s 70 -c 16 -o $O/r02_enc $E > $O/r02_prof_enc_ncu.log 2>&1
export_rep r02_enc source
for EH in 8 128; do
H="python tools/time_head.py --episodes $EH --iters 2"
timeout 300 $H > $O/r02_prof_head${EH}_plain.log 2>&1 &&
timeout 600 ncu $LIGHT --clock-control none -s $((3*42)) -c 42 -o $O/r02_head_e$EH $H > $O/r02_prof_head${EH}_ncu.log 2>&1
export_rep r02_head_e$EH
done
T="python tools/time_head_kernels.py --one"
timeout 300 $T > $O/r02_prof_otam_plain.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"cos_dist|otam" -s 3 -c 2 -o $O/r02_otam $T > $O/r02_prof_otam_ncu.log 2>&1
export_rep r02_otam source
T4="python tools/time_head_kernels.py --p4000"
timeout 300 $T4 > $O/r02_prof_otam4k_plain.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"cos_dist|otam" -s 3 -c 2 -o $O/r02_otam4k $T4 > $O/r02_prof_otam4k_ncu.log 2>&1
export_rep r02_otam4k source
R="python tools/rn50_throughput.py"
timeout 300 $R > $O/r02_prof_rn50_plain.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 800 -c 900 --csv --log-file $O/r02_launches_rn50.csv $R > $O/r02_prof_rn50_ncu.log 2>&1
for f in $O/r02_prof_*_plain.log; do echo "== $f"; tail -n 2 $f; done
du -sh $O; ls -la $O | tail -30
