#!/bin/bash
mkdir -p gpurun_out
W=${WORKLOAD:-c2}
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"${KERNELS:-k_dust|k_score|k_select}" -s ${SKIP:-4} -c ${COUNT:-3} -o gpurun_out/prof5_${W} -f python bench.py --workload $W --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu5.log 2>&1
tail -1 gpurun_out/ncu5.log
