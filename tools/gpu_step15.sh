#!/bin/bash
# LayerNorm folded into the encoder GEMM epilogues: parity, then interleaved A/B of the bench step
set -x
O=gpurun_out
timeout 900 python -m pytest tests/test_stages_gpu.py tests/test_canaries_gpu.py tests/test_frames_gpu.py -x -q > $O/r02_lnf_tests.log 2>&1; tail -n 12 $O/r02_lnf_tests.log
B="python bench.py --steps 6 --warmup 3 --no-extra-legs --no-cpu-baseline"
for r in 1 2; do
  SPM_LN_FOLD=0 timeout 300 $B > $O/r02_lnf_off_$r.json 2> $O/r02_lnf_off_$r.err
  SPM_LN_FOLD=1 timeout 300 $B > $O/r02_lnf_on_$r.json 2> $O/r02_lnf_on_$r.err
done
for f in $O/r02_lnf_o*_?.json; do echo "== $f"; python - "$f" <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d.get("e2e",{}).get("value"), d["roofline"]["frac"], d.get("clocks"))
PY
done
