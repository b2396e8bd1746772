#!/bin/bash
O=gpurun_out
echo "== rn50 tests, window conv with base offset"; timeout 600 python -m pytest tests -m gpu -x -q -k "rn50" 2>&1 | tail -n 4
echo "== rn50 tests, window conv WITHOUT base offset"; SPM_CONV_WIN_BO=0 timeout 600 python -m pytest tests -m gpu -x -q -k "rn50" 2>&1 | tail -n 4
echo "== rn50 tests, window conv off"; SPM_CONV_WIN=0 timeout 600 python -m pytest tests -m gpu -x -q -k "rn50" 2>&1 | tail -n 3
( timeout 300 python tools/rn50_throughput.py 8 6; SPM_CONV_WIN_BO=0 timeout 300 python tools/rn50_throughput.py 8 6; SPM_CONV_WIN=0 timeout 300 python tools/rn50_throughput.py 8 6 ) 2>&1 | grep -v "^+" | cut -c40-
echo "== jpeg"; timeout 600 python -m pytest tests -m gpu -x -q -k "jpeg" 2>&1 | tail -n 12
