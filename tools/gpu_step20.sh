#!/bin/bash
set -x
O=gpurun_out
cat > /tmp/otf.py <<'PY'
import torch, sys
sys.path.insert(0, ".")
from clip_spm_b200 import ops
sup = torch.randn(300, 5, 8, 512, device="cuda"); tgt = torch.randn(300, 5, 8, 512, device="cuda")
out = ops.otam_distance(sup, tgt, False); torch.cuda.synchronize(); print(out[0, 0])
PY
SPM_OTAM_FUSED=1 SPM_OTAM_PF=1 timeout 300 compute-sanitizer --tool memcheck python /tmp/otf.py > $O/r02_s20_sanitizer.log 2>&1
grep -v "^=========     Host Frame\|^=========         in " $O/r02_s20_sanitizer.log | head -n 40
T="python tools/time_head_kernels.py --one"
SPM_OTAM_PF=0 SPM_OTAM_FUSED=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"otam_fused" -s 3 -c 1 -o $O/r02_s19_otamf $T > $O/r02_s19_ncu.log 2>&1
ncu -i $O/r02_s19_otamf.ncu-rep --page raw --csv > $O/r02_s19_otamf_raw.csv 2>/dev/null
ncu -i $O/r02_s19_otamf.ncu-rep --page source --csv > $O/r02_s19_otamf_source.csv 2>/dev/null
rm -f $O/r02_s19_otamf.ncu-rep
tail -n 3 $O/r02_s19_ncu.log
