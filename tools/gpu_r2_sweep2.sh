#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python tools/sweep.py c4 3 CFR_B200_PAIR_SEARCH_BLOCKS=5 CFR_B200_PAIR_SEARCH_BLOCKS=5,CFR_B200_DUST_LANES=32 CFR_B200_PAIR_SEARCH_BLOCKS=5,CFR_B200_DUST_QUORUM=4 CFR_B200_PAIR_SEARCH_BLOCKS=5,CFR_B200_DUST_QUORUM=16 CFR_B200_PAIR_SEARCH_BLOCKS=5,CFR_B200_QUORUM=12 > gpurun_out/sweep2_c4.jsonl 2> gpurun_out/sweep2_c4.err
tail -2 gpurun_out/sweep2_c4.err; cat gpurun_out/sweep2_c4.jsonl
SWEEP_STREAMS=4 timeout 900 python tools/sweep.py c4 4 CFR_B200_PAIR_SEARCH_BLOCKS=5 > gpurun_out/sweep2_c4_s4.jsonl 2>> gpurun_out/sweep2_c4.err
cat gpurun_out/sweep2_c4_s4.jsonl
SWEEP_STREAMS=2 timeout 900 python tools/sweep.py c4 4 CFR_B200_PAIR_SEARCH_BLOCKS=5 > gpurun_out/sweep2_c4_s2.jsonl 2>> gpurun_out/sweep2_c4.err
cat gpurun_out/sweep2_c4_s2.jsonl
