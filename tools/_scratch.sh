timeout 600 python -m pytest tests/test_stages_gpu.py tests/test_frames_gpu.py -x -q -k "rn50" 2>&1 | tail -3
timeout 300 python tools/rn50_throughput.py 2>&1 | tail -1
SPM_CONV_PAIR=0 timeout 300 python tools/rn50_throughput.py 2>&1 | tail -1
