#!/bin/bash
# (1) --set full of the final encoder kernels (skip count for the LayerNorm-folded launch sequence); (2) tcgen05 OTAM with the
# A operands in tensor memory: parity + timing; (3) JPEG decode timing
set -x
O=gpurun_out
E="python tools/profile_encoder.py 512 3"
timeout 300 $E > $O/r02_prof_enc_plain.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"gemm2_tcgen05|vit_attention_tc|layernorm" -s 70 -c 16 -o $O/r02_enc $E > $O/r02_prof_enc_ncu.log 2>&1
ncu -i $O/r02_enc.ncu-rep --page raw --csv > $O/r02_enc_raw.csv 2>/dev/null
ncu -i $O/r02_enc.ncu-rep --page source --csv > $O/r02_enc_source.csv 2>/dev/null
rm -f $O/r02_enc.ncu-rep
SPM_OTAM_TC_MINP=296 timeout 300 python tools/otam_dp_check.py > $O/r02_s23_check.log 2>&1; grep "P=300 W=5 Q=5 T=8\|P=1974\|P=1800\|worst" $O/r02_s23_check.log | cut -c1-150
SPM_OTAM_TC_MINP=296 timeout 300 python tools/time_head_kernels.py 2>&1 | tail -n 4 > $O/r02_s23_head_kernels.log 2>&1
cat $O/r02_s23_head_kernels.log
timeout 300 python tools/time_jpeg.py > $O/r02_s23_jpeg.log 2>&1; tail -n 2 $O/r02_s23_jpeg.log
