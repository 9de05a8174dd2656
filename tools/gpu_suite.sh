#!/bin/bash
# the GPU suite + smoke, as the driver runs them
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu_last.log 2>&1
tail -4 gpurun_out/pytest_gpu_last.log
( timeout 300 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
