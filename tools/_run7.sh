timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
timeout 600 python bench.py > gpurun_out/bench_a.json 2> gpurun_out/bench_a.err; tail -2 gpurun_out/bench_a.err; cat gpurun_out/bench_a.json
timeout 300 python bench.py --no-cpu-baseline --no-e2e --steps 16 2>/dev/null | tail -1 | cut -c1-200
