#!/bin/bash
set -x
O=gpurun_out
R="python tools/rn50_throughput.py 2 3"
timeout 300 $R > $O/r02_s6_plain.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"gemm_tcgen05_kernel|gemm2_tcgen05" -s 210 -c 9 -o $O/r02_rn50_front $R > $O/r02_s6_ncu.log 2>&1
ncu -i $O/r02_rn50_front.ncu-rep --page raw --csv > $O/r02_rn50_front_raw.csv 2>/dev/null
ncu -i $O/r02_rn50_front.ncu-rep --page source --csv > $O/r02_rn50_front_source.csv 2>/dev/null
rm -f $O/r02_rn50_front.ncu-rep
cat $O/r02_s6_plain.log
( SPM_PDL=0 timeout 300 python tools/rn50_throughput.py 8 6; timeout 300 python tools/rn50_throughput.py 8 6; SPM_RN50_FRONT_CHUNK=108 timeout 300 python tools/rn50_throughput.py 8 6 ) 2>&1 | grep -v "^+" > $O/r02_rn50_pdl.log; cat $O/r02_rn50_pdl.log
timeout 900 python -m pytest tests -m gpu -x -q -k "rn50 or gemm or vit" > $O/r02_s6_tests.log 2>&1; tail -n 4 $O/r02_s6_tests.log
( SPM_PDL=0 timeout 300 python bench.py --steps 6 --warmup 3 --no-e2e --no-extra-legs --no-cpu-baseline; timeout 300 python bench.py --steps 6 --warmup 3 --no-e2e --no-extra-legs --no-cpu-baseline ) 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('bench value', d['value'], 'gemm frac', d['roofline']['frac'], 'step_ms', d['step_ms'])" > $O/r02_bench_pdl.log; cat $O/r02_bench_pdl.log
du -sh $O
