run() { echo "== $*"; env "$@" timeout 300 python bench.py --steps 16 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('value %.2f step_ms %s gemm %.0f TF/s  clocks %s' % (d['value'], d['step_ms'], d['roofline']['achieved'], d['clocks']['sm_mhz']))"; }
for i in 1 2; do
run SPM_GEMM_2CTA=1
run SPM_GEMM_2CTA=2
run SPM_GEMM_2CTA=2 SPM_FRAME_CHUNK=512
run SPM_GEMM_2CTA=2 SPM_FRAME_CHUNK=384
run SPM_GEMM_2CTA=2 SPM_FRAME_CHUNK=960
done
