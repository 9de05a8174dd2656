#!/bin/bash
# one-process sweep of CFR_B200_* settings (space-separated in $SWEEP_SETTINGS, switches of one setting joined by commas)
mkdir -p gpurun_out
timeout 1500 python tools/sweep.py ${SWEEP_WORKLOAD:-c4} ${SWEEP_BATCHES:-3} $SWEEP_SETTINGS > gpurun_out/sweep_last.jsonl 2> gpurun_out/sweep_last.err
tail -2 gpurun_out/sweep_last.err; cat gpurun_out/sweep_last.jsonl
