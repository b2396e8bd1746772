run() { echo "== $*"; env "$@" timeout 300 python bench.py --steps 16 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('value %.2f median %.1f max %.1f gemm %.0f clocks %s' % (d['value'], d['step_ms']['median'], d['step_ms']['max'], d['roofline']['achieved'], d['clocks']['sm_mhz']))"; }
timeout 600 python -m pytest tests/test_stages_gpu.py tests/test_gemm_gpu.py -x -q 2>&1 | tail -2
for i in 1 2; do
run SPM_ALT_DIR=1
run SPM_ALT_DIR=0
done
