#!/bin/bash
# quick GPU iteration: parity tests, then bench variants. usage: gpu_iter.sh "<workloads>" "<variants>"
# variants: occ | rb | qN (quorum N) | bN (search blocks/SM N)
mkdir -p gpurun_out
WL=${1:-"c2"}
VARS=${2:-"occ rb"}
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
grep -E "passed|failed|rror" gpurun_out/pytest_gpu.log | tail -3
for W in $WL; do
  for V in $VARS; do
    unset CFR_B200_QUORUM CFR_B200_SEARCH_BLOCKS
    if [[ $V == q* ]]; then export CFR_B200_QUORUM=${V#q}; fi
    if [[ $V == b* ]]; then export CFR_B200_SEARCH_BLOCKS=${V#b}; fi
    EXTRA=""
    if [ $V = rb ]; then EXTRA="--layout 1"; fi
    timeout 600 python bench.py --workload $W --steps 5 --warmup 3 --no-cpu-baseline $EXTRA > gpurun_out/bench_${W}_${V}.json 2> gpurun_out/bench_${W}_${V}.err
    python - <<PY
import json
f="gpurun_out/bench_${W}_${V}.json"
try:
    d=json.loads(open(f).read().strip().splitlines()[-1])
    print("${W} ${V}: value %.4g e2e %.4g ms/step %.3f"%(d["value"], d["e2e"]["value"], d["ms_per_step"]), {k:round(v,3) for k,v in d["stage_ms_per_step"].items()}, "search GB/s %.0f"%d["roofline"]["achieved"], d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
except Exception as e:
    print(f, "FAILED", e); print(open(f.replace(".json",".err")).read()[-1500:])
PY
  done
done
