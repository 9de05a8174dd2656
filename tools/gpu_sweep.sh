#!/bin/bash
mkdir -p gpurun_out
export CFR_B200_SCALAR_OCC=1
for Q in 1 2 4 16 32; do
  CFR_B200_QUORUM=$Q timeout 300 python bench.py --workload c2 --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/bench_q$Q.json 2> gpurun_out/bench_q$Q.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/bench_q$Q.json").read().strip().splitlines()[-1])
print("quorum $Q:", {k:round(v,3) for k,v in d["stage_ms_per_step"].items()})
PY
done
CFR_B200_QUORUM=8 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_search|k_locate" -s 2 -c 2 -o gpurun_out/prof3_c2_scalar -f python bench.py --workload c2 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu3.log 2>&1
tail -1 gpurun_out/ncu3.log
