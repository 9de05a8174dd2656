#!/bin/bash
# first contact with the GPU: parity tests, smoke, bench (both layouts), launch list
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
tail -15 gpurun_out/pytest_gpu.log
( time timeout 300 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/smoke.log 2>&1
tail -3 gpurun_out/smoke.log
timeout 600 python bench.py --workload c2 --steps 3 --warmup 3 > gpurun_out/bench_c2_occ.json 2> gpurun_out/bench_c2_occ.err
tail -c 3000 gpurun_out/bench_c2_occ.json; tail -5 gpurun_out/bench_c2_occ.err
timeout 600 python bench.py --workload c2 --steps 3 --warmup 3 --layout 1 --no-cpu-baseline > gpurun_out/bench_c2_rb.json 2> gpurun_out/bench_c2_rb.err
tail -c 3000 gpurun_out/bench_c2_rb.json; tail -5 gpurun_out/bench_c2_rb.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_c2.csv python bench.py --workload c2 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
tail -3 gpurun_out/ncu_bench.log
