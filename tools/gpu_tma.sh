#!/bin/bash
# the bulk-copy (TMA) staging variant of the search round: parity first (short timeouts: a wrong barrier protocol hangs), then A/B
mkdir -p gpurun_out
timeout 240 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "variants and (tma or default)" > gpurun_out/pytest_tma.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/pytest_tma.log
if grep -q "passed" gpurun_out/pytest_tma.log && ! grep -q "failed" gpurun_out/pytest_tma.log; then
  SWEEP_STEPS=3 timeout 420 python tools/sweep.py c4 3 default CFR_B200_PAIR_FETCH=4 CFR_B200_PAIR_FETCH=4,CFR_B200_PAIR_SEARCH_BLOCKS=8 > gpurun_out/sweep_tma.jsonl 2> gpurun_out/sweep_tma.err
  echo "sweep rc=$?"; tail -2 gpurun_out/sweep_tma.err; cat gpurun_out/sweep_tma.jsonl
fi
