#!/bin/bash
# the whole GPU suite + smoke, as the driver runs them
mkdir -p gpurun_out
( time timeout 2400 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu_full.log 2>&1
tail -6 gpurun_out/pytest_gpu_full.log
( time timeout 600 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/smoke.log 2>&1; tail -4 gpurun_out/smoke.log
