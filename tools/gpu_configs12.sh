#!/bin/bash
# BASELINE configs[1] (100 Mbp index, 1 M x 100 bp single-end) and configs[2] (2 Gbp index from the unmodified reference builder,
# 10 M x 2x150 bp pairs, -k 5) with the current code, one bench line each
mkdir -p gpurun_out
( time timeout 900 python bench.py --workload c2 --steps 50 --warmup 5 ) > gpurun_out/r02_bench_c2.json 2> gpurun_out/r02_bench_c2.err
tail -2 gpurun_out/r02_bench_c2.err; cut -c1-300 gpurun_out/r02_bench_c2.json
( time timeout 1800 python bench.py --workload c3 --steps 10 --warmup 3 ) > gpurun_out/r02_bench_c3.json 2> gpurun_out/r02_bench_c3.err
tail -3 gpurun_out/r02_bench_c3.err; cut -c1-300 gpurun_out/r02_bench_c3.json
