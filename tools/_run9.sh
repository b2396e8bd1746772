set -e
timeout 300 python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/b_pre2.json 2>gpurun_out/b_pre2.err
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:gemm2_tcgen05 -s 60 -c 8 -f -o gpurun_out/gemm2_full python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out/gemm2_full.ncu-rep
