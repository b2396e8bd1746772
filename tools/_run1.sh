set -x
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -5
timeout 300 python tools/time_frame_transform.py 2>&1 | tail -8
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_u8.json 2> gpurun_out/bench_u8.err; tail -3 gpurun_out/bench_u8.err; cat gpurun_out/bench_u8.json
