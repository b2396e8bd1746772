timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_frames_gpu.py tests/test_otam_backward_gpu.py -x -q -k "transform_frames or otam_backward_matches_autograd or golden" 2>&1 | tail -15
echo "exit: $?"
