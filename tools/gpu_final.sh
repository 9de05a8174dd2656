#!/bin/bash
# closing pass of a round, as the driver runs things: GPU suite, smoke, bench (both arms), ncu captures for the roofline fields
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu_final.log 2>&1
tail -3 gpurun_out/pytest_gpu_final.log
( time timeout 600 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/smoke.log 2>&1; tail -4 gpurun_out/smoke.log
( time timeout 1500 python bench.py ) > gpurun_out/r02_bench_default.json 2> gpurun_out/r02_bench_default.err
tail -3 gpurun_out/r02_bench_default.err; cut -c1-400 gpurun_out/r02_bench_default.json
( time timeout 1500 python bench.py --impl reference ) > gpurun_out/r02_bench_reference.json 2> gpurun_out/r02_bench_reference.err
tail -3 gpurun_out/r02_bench_reference.err; cut -c1-300 gpurun_out/r02_bench_reference.json
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_search" -s 2 -c 1 -o gpurun_out/prof_r02_c4_final -f python bench.py --reads 1000000 --steps 1 --warmup 1 --no-cpu-baseline --no-cli > gpurun_out/ncu_full_c4.log 2>&1
tail -2 gpurun_out/ncu_full_c4.log
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 120 --csv --log-file gpurun_out/r02_launches_c4.csv python bench.py --reads 1000000 --steps 1 --warmup 1 --no-cpu-baseline --no-cli > gpurun_out/ncu_launch_c4.log 2>&1
grep -c "k_" gpurun_out/r02_launches_c4.csv
