#!/bin/bash
# ncu --set full with source counters of the tcgen05 OTAM kernel (P = 4000: steady state)
set -x
O=gpurun_out
cat > /tmp/one.py <<'PY'
import torch, sys
sys.path.insert(0, ".")
from clip_spm_b200 import ops
P = 4000
sup = torch.randn(P, 5, 8, 512, device="cuda"); tgt = torch.randn(P, 5, 8, 512, device="cuda")
out = torch.zeros(P, 5, 5, device="cuda")
for i in range(3):
    ops.otam_distance(sup, tgt, False, out=out)
torch.cuda.synchronize(); print(float(out.sum()))
PY
timeout 120 python /tmp/one.py
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"otam_tc" -s 2 -c 1 -o $O/r02_s22_otamtc python /tmp/one.py > $O/r02_s22_ncu.log 2>&1
ncu -i $O/r02_s22_otamtc.ncu-rep --page raw --csv > $O/r02_s22_otamtc_raw.csv 2>/dev/null
ncu -i $O/r02_s22_otamtc.ncu-rep --page source --csv > $O/r02_s22_otamtc_source.csv 2>/dev/null
rm -f $O/r02_s22_otamtc.ncu-rep
tail -n 3 $O/r02_s22_ncu.log
