#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_multi.py -q -m gpu -x -p no:cacheprovider > gpurun_out/pytest_multi.log 2>&1
echo "pytest_multi rc=$?"; tail -5 gpurun_out/pytest_multi.log; grep -n "timeout tag" gpurun_out/pytest_multi.log | head -5
