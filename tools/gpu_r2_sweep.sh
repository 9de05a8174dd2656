#!/bin/bash
# GPU suite, then a one-process sweep of kernel switches on the 20 Gbp workload
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu_sweep.log 2>&1
tail -4 gpurun_out/pytest_gpu_sweep.log
timeout 1500 python tools/sweep.py c4 3 ${SWEEP_SETTINGS:-"" CFR_B200_PAIR_FETCH=2 CFR_B200_PAIR_FETCH=1 CFR_B200_PAIR_SEARCH_BLOCKS=8 CFR_B200_PAIR_SEARCH_BLOCKS=7 CFR_B200_PAIR_SEARCH_BLOCKS=5 CFR_B200_PAIR_FETCH=2,CFR_B200_PAIR_SEARCH_BLOCKS=8 CFR_B200_DENSE_LOCATE=1 CFR_B200_QUORUM=6 CFR_B200_QUORUM=12} \
  > gpurun_out/sweep_c4.jsonl 2> gpurun_out/sweep_c4.err
tail -3 gpurun_out/sweep_c4.err
cat gpurun_out/sweep_c4.jsonl
