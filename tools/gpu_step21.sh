#!/bin/bash
# fused OTAM kernel: parity + batch-scale timing (two-kernel default vs fused, 32- and 16-column ring stages)
set -x
O=gpurun_out
SPM_OTAM_KC=16 timeout 300 python tools/otam_dp_check.py > $O/r02_s21_check_kc16.log 2>&1; tail -n 2 $O/r02_s21_check_kc16.log
SPM_OTAM_FUSED=1 SPM_OTAM_KC=16 timeout 300 python tools/otam_dp_check.py > $O/r02_s21_check_kc16.log 2>&1; tail -n 2 $O/r02_s21_check_kc16.log
timeout 900 python -m pytest tests/test_stages_gpu.py tests/test_canaries_gpu.py -q -k "otam or canar" > $O/r02_s21_tests.log 2>&1; tail -n 8 $O/r02_s21_tests.log
for v in "SPM_OTAM_FUSED=1" "SPM_OTAM_FUSED=1 SPM_OTAM_KC=16"; do
  echo "== [$v]"; env $v timeout 300 python tools/time_head_kernels.py 2>&1 | tail -n 4
done > $O/r02_s21_head_kernels.log 2>&1
cat $O/r02_s21_head_kernels.log
