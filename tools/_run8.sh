set -e
timeout 300 python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/b_pre.json 2>gpurun_out/b_pre.err
tail -c 300 gpurun_out/b_pre.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_final.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_b.log 2>&1
wc -l gpurun_out/launches_final.csv
