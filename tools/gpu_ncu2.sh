#!/bin/bash
mkdir -p gpurun_out
W=${WORKLOAD:-c2}
for V in ${VARIANTS:-scalar}; do
  if [ $V = scalar ]; then export CFR_B200_SCALAR_OCC=1; else unset CFR_B200_SCALAR_OCC; fi
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_search|k_locate|k_dust|k_encode" -s 4 -c 4 \
     -o gpurun_out/prof2_${W}_${V} -f python bench.py --workload $W --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu2_${V}.log 2>&1
  tail -2 gpurun_out/ncu2_${V}.log
done
