#!/bin/bash
set -x
O=gpurun_out
timeout 300 python tools/rn50_shapes.py > $O/r02_s24_rn50_shapes.log 2>&1; sort $O/r02_s24_rn50_shapes.log | tail -n 40
timeout 900 python bench.py > $O/r02_s24_bench.json 2> $O/r02_s24_bench.err; tail -c 3000 $O/r02_s24_bench.json
