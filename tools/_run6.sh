run() { echo "== $*"; env "$@" timeout 300 python bench.py --steps 16 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('value %.2f median %.1f max %.1f clocks %s samples %s' % (d['value'], d['step_ms']['median'], d['step_ms']['max'], d['clocks']['sm_mhz'], d['clocks']['samples']))"; }
for i in 1 2 3; do
run SPM_GEMM_2CTA=2 SPM_FRAME_CHUNK=512 SPM_BENCH_CLOCK_INTERVAL=0.4
run SPM_GEMM_2CTA=2 SPM_FRAME_CHUNK=512 SPM_BENCH_CLOCK_INTERVAL=5
run SPM_GEMM_2CTA=2 SPM_FRAME_CHUNK=512 SPM_ENC_STREAMS=1 SPM_BENCH_CLOCK_INTERVAL=0.4
done
uptime; nproc
