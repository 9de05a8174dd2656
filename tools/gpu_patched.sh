#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "patched" > gpurun_out/pytest_patched.log 2>&1
echo "rc=$?"; tail -15 gpurun_out/pytest_patched.log
