#!/bin/bash
set -x
O=gpurun_out
for d in 0 1 2 3; do echo "== dbg $d"; SPM_OTAM_TC_DBG=$d timeout 200 python tools/time_head_kernels.py --p4000 2>&1 | tail -n 1; done > $O/r02_s27_dbg.log 2>&1
cat $O/r02_s27_dbg.log
for n in 960 3840 7680; do timeout 300 python tools/time_jpeg.py $n 2>&1 | tail -n 1; done > $O/r02_s27_jpeg.log 2>&1
cat $O/r02_s27_jpeg.log
