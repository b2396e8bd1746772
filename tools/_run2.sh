run() { echo "== $*"; env "$@" timeout 300 python bench.py --steps 8 --warmup 3 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('value %.2f  step_ms %s  gemm %.0f TF/s  clocks %s' % (d['value'], d['step_ms'], d['roofline']['achieved'], d['clocks']['sm_mhz']))"; }
timeout 600 python -m pytest tests/test_stages_gpu.py -x -q 2>&1 | tail -3
run SPM_ENC_STREAMS=1 SPM_FRAME_CHUNK=256
run SPM_ENC_STREAMS=2 SPM_FRAME_CHUNK=256
run SPM_ENC_STREAMS=2 SPM_FRAME_CHUNK=128
run SPM_ENC_STREAMS=2 SPM_FRAME_CHUNK=126
run SPM_ENC_STREAMS=2 SPM_FRAME_CHUNK=63
run SPM_ENC_STREAMS=1 SPM_FRAME_CHUNK=63
run SPM_ENC_STREAMS=2 SPM_FRAME_CHUNK=192
