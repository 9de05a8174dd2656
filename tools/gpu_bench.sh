#!/bin/bash
# GPU suite, the driver's two bench invocations (default workload), and the ncu captures the roofline fields cite
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q --durations=8 ) > gpurun_out/pytest_gpu_bench.log 2>&1
tail -14 gpurun_out/pytest_gpu_bench.log
( time timeout 1500 python bench.py ) > gpurun_out/r02_bench_default.json 2> gpurun_out/r02_bench_default.err
tail -3 gpurun_out/r02_bench_default.err; cut -c1-6000 gpurun_out/r02_bench_default.json
if [ -z "$SKIP_REF" ]; then
( time timeout 1500 python bench.py --impl reference ) > gpurun_out/r02_bench_reference.json 2> gpurun_out/r02_bench_reference.err
tail -3 gpurun_out/r02_bench_reference.err; cut -c1-1500 gpurun_out/r02_bench_reference.json
fi
if [ -z "$SKIP_NCU" ]; then
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_search" -s 2 -c 1 -o gpurun_out/prof_r02_c4_single_wait -f python bench.py --reads 1000000 --steps 1 --warmup 1 --no-cpu-baseline --no-cli > gpurun_out/ncu_full_c4.log 2>&1
tail -2 gpurun_out/ncu_full_c4.log
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 80 --csv --log-file gpurun_out/r02_launches_c4.csv python bench.py --reads 1000000 --steps 1 --warmup 1 --no-cpu-baseline --no-cli > gpurun_out/ncu_launch_c4.log 2>&1
grep -c "k_" gpurun_out/r02_launches_c4.csv
fi
if [ -n "$SWEEP_SETTINGS" ]; then
timeout 1500 python tools/sweep.py c4 3 $SWEEP_SETTINGS > gpurun_out/sweep_c4.jsonl 2> gpurun_out/sweep_c4.err
tail -3 gpurun_out/sweep_c4.err; cat gpurun_out/sweep_c4.jsonl
fi
