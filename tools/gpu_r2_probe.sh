#!/bin/bash
# round 2, first contact: box facts, c3 (BASELINE configs[2]) baseline bench + ncu capture of the search kernel
mkdir -p gpurun_out
{ nproc; free -g; df -h /tmp /dev/shm . ; nvidia-smi --query-gpu=name,memory.total,memory.used --format=csv; lscpu | head -20; ulimit -a; } > gpurun_out/box_facts.txt 2>&1
( time python -c "
import sys; sys.path.insert(0,'tools')
import make_data
print(make_data.ensure('c3'))" ) > gpurun_out/build_c3.log 2>&1
tail -4 gpurun_out/build_c3.log
ls -la data/c3 >> gpurun_out/box_facts.txt
timeout 900 python bench.py --workload c3 --steps 10 --warmup 3 > gpurun_out/r02_bench_c3_base.json 2> gpurun_out/r02_bench_c3_base.err
cut -c1-1500 gpurun_out/r02_bench_c3_base.json
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_search" -s 1 -c 1 -o gpurun_out/prof_r02_c3_base -f python bench.py --workload c3 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full_c3.log 2>&1
tail -2 gpurun_out/ncu_full_c3.log
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 40 --csv --log-file gpurun_out/r02_launches_c3_base.csv python bench.py --workload c3 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launch_c3.log 2>&1
grep -c k_ gpurun_out/r02_launches_c3_base.csv
