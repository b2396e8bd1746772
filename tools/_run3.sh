run() { echo "== $*"; env "$@" timeout 300 python bench.py --steps 12 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('value %.2f  e2e %.2f  e2e_u8 %.2f step_ms %s prof %s gemm %.0f TF/s  clocks %s' % (d['value'], d['e2e']['value'], d['e2e_u8']['value'], d['step_ms'], d['roofline']['step_ms_profiled'], d['roofline']['achieved'], d['clocks']['sm_mhz']))"; }
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
run SPM_ENC_STREAMS=2
run SPM_ENC_STREAMS=1
run SPM_ENC_STREAMS=2 SPM_FRAME_CHUNK=320
run SPM_ENC_STREAMS=2 SPM_FRAME_CHUNK=240
