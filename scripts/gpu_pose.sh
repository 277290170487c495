#!/bin/bash
# pose/corresp kernels: sanitizer pass on the small tests, then the full pose test file
mkdir -p gpurun_out
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 99 python -m pytest tests/test_pose_gpu.py -x -q -k "corresp_planted or dropin or small_and_degenerate or golden" > gpurun_out/sanitizer_pose.log 2>&1
echo "sanitizer rc=$?"; tail -25 gpurun_out/sanitizer_pose.log
timeout 900 python -m pytest tests/test_pose_gpu.py -x -q 2>&1 | tail -40 | tee gpurun_out/pytest_pose.log
