#!/bin/bash
# exponent-domain OTAM wavefront: parity in every kernel, batch-scale timings (two-kernel, fused, log-domain)
set -x
O=gpurun_out
timeout 900 python -m pytest tests/test_stages_gpu.py tests/test_canaries_gpu.py -q -k "otam or canar" > $O/r02_s18_tests.log 2>&1; tail -n 25 $O/r02_s18_tests.log
for v in "" "SPM_OTAM_DP=log" "SPM_OTAM_FUSED=1" "SPM_OTAM_FUSED=1 SPM_OTAM_DP=log"; do
  echo "== [$v]"; env $v timeout 300 python tools/time_head_kernels.py 2>&1 | tail -n 6
done > $O/r02_s18_head_kernels.log 2>&1
cat $O/r02_s18_head_kernels.log
