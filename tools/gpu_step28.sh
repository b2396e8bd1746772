#!/bin/bash
set -x
O=gpurun_out
SPM_OTAM_TC_MINP=296 timeout 300 python tools/otam_dp_check.py > $O/r02_s28_check.log 2>&1; grep "worst" $O/r02_s28_check.log
SPM_OTAM_TC_FLAT=0 SPM_OTAM_TC_MINP=296 timeout 300 python tools/otam_dp_check.py > $O/r02_s28_check4d.log 2>&1; grep "worst" $O/r02_s28_check4d.log
for v in "SPM_OTAM_TC_MINP=296" "SPM_OTAM_TC_MINP=296 SPM_OTAM_TC_DBG=3" "SPM_OTAM_TC_MINP=296 SPM_OTAM_TC_FLAT=0"; do echo "== $v"; env $v timeout 200 python tools/time_head_kernels.py 2>&1 | tail -n 4; done > $O/r02_s28_times.log 2>&1
grep -v "^+" $O/r02_s28_times.log
timeout 600 python -m pytest tests/test_optim_gpu.py -q > $O/r02_s28_optim.log 2>&1; tail -n 3 $O/r02_s28_optim.log
