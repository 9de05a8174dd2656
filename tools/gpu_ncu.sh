#!/bin/bash
# one ncu --set full capture of the pipeline kernels (first timed step), scalar and cooperative occ variants
mkdir -p gpurun_out
W=${WORKLOAD:-c2}
for V in coop scalar; do
  if [ $V = scalar ]; then export CFR_B200_SCALAR_OCC=1; else unset CFR_B200_SCALAR_OCC; fi
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_search|k_locate|k_dust|k_score|k_select" -s 5 -c 5 \
     -o gpurun_out/prof_${W}_${V} -f python bench.py --workload $W --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_${V}.log 2>&1
  tail -2 gpurun_out/ncu_${V}.log
done
ls -la gpurun_out/*.ncu-rep
