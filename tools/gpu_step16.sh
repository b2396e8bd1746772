#!/bin/bash
# LayerNorm folding variants: feature error vs the reference golden, per-shape GEMM times in situ, step A/B
set -x
O=gpurun_out
for m in 0 1 2; do SPM_LN_FOLD=$m timeout 300 python tools/lnf_error.py >> $O/r02_lnf_error.log 2>&1; done
timeout 300 python tools/lnf_error.py vit_5w5s_t8_p1 bf16_resid >> $O/r02_lnf_error.log 2>&1
cat $O/r02_lnf_error.log
B="python bench.py --steps 6 --warmup 3 --no-extra-legs --no-cpu-baseline --no-e2e"
for r in 1 2; do for m in 0 1 2; do
  SPM_PROFILE_SHAPES=1 SPM_LN_FOLD=$m timeout 300 $B > $O/r02_lnf_m${m}_$r.json 2> $O/r02_lnf_m${m}_$r.err
done; done
for f in $O/r02_lnf_m?_?.json; do echo "== $f"; python - "$f" <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["roofline"]["frac"], d["roofline"]["gemm_share_of_step"], d.get("clocks"))
PY
grep "spm profile" ${f%.json}.err | sort | uniq | head -12
done
timeout 900 python -m pytest tests/test_stages_gpu.py tests/test_canaries_gpu.py tests/test_frames_gpu.py tests/test_sweep_gpu.py -q > $O/r02_lnf_tests.log 2>&1; tail -n 12 $O/r02_lnf_tests.log
