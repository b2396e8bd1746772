#!/bin/bash
# final r02 evidence: full GPU suite, bench lines (own arm + reference arm), then the ncu passes of tools/gpu_profile_r02.sh
set -x
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > $O/r02_final_tests.log 2>&1; tail -n 6 $O/r02_final_tests.log
timeout 600 python bench.py --impl reference > $O/r02_final_bench_reference.json 2> $O/r02_final_bench_reference.err; tail -c 600 $O/r02_final_bench_reference.json
timeout 900 python bench.py > $O/r02_final_bench.json 2> $O/r02_final_bench.err; tail -c 400 $O/r02_final_bench.json
bash tools/gpu_profile_r02.sh > $O/r02_final_profile.log 2>&1; tail -n 25 $O/r02_final_profile.log
