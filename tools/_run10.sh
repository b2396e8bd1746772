timeout 600 python -m pytest tests/test_gemm_gpu.py -x -q 2>&1 | tail -5
timeout 300 python tools/time_residual_gemm.py 2>&1 | tail -7
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
for i in 1 2; do
for v in 1 0; do echo "RES_TMA=$v"; SPM_GEMM_RES_TMA=$v timeout 300 python bench.py --steps 16 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('value %.2f median %.1f max %.1f gemm %.0f clocks %s' % (d['value'], d['step_ms']['median'], d['step_ms']['max'], d['roofline']['achieved'], d['clocks']['sm_mhz']))"; done; done
