#!/bin/bash
O=gpurun_out
( for f in 27 54 72 108; do SPM_RN50_FRONT_CHUNK=$f SPM_RN50_BACK_CHUNK=216 timeout 300 python tools/rn50_throughput.py 8 6; done
  SPM_RN50_FRONT_CHUNK=64 SPM_RN50_BACK_CHUNK=320 timeout 300 python tools/rn50_throughput.py 8 6
  SPM_RN50_FRONT_CHUNK=320 SPM_RN50_BACK_CHUNK=320 timeout 300 python tools/rn50_throughput.py 8 6 ) 2>&1 | grep -v "^+" > $O/r02_rn50_s10.log; cat $O/r02_rn50_s10.log
