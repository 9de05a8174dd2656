#!/bin/bash
# quick check: parity tests + bench variants. usage: gpu_quick.sh "<workloads>" "<variants>"  (m700 is built on the box)
mkdir -p gpurun_out
WL=${1:-"c2"}
VARS=${2:-"default"}
if [[ "$WL" == *m700* ]]; then
  ( python -c "
import sys; sys.path.insert(0,'tools')
import make_data
make_data.ensure('m700')" > gpurun_out/m700_build.log 2>&1 ) &
  BUILD_PID=$!
fi
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
grep -E "passed|failed|rror" gpurun_out/pytest_gpu.log | tail -3
for W in $WL; do
  if [[ $W == m700* ]] && [ -n "$BUILD_PID" ]; then wait $BUILD_PID; BUILD_PID=""; fi
  for V in $VARS; do
    unset CFR_B200_QUORUM CFR_B200_SEARCH_BLOCKS CFR_B200_DUST_SCREEN CFR_B200_L2_FETCH
    case $V in
      noscreen) export CFR_B200_DUST_SCREEN=0 ;;
      sb*) export CFR_B200_SEARCH_BLOCKS=${V#sb} ;;
      l2*) export CFR_B200_L2_FETCH=${V#l2} ;;
      q*) export CFR_B200_QUORUM=${V#q} ;;
    esac
    timeout 600 python bench.py --workload $W --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${W}_${V}.json 2> gpurun_out/bench_${W}_${V}.err
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
unset CFR_B200_QUORUM CFR_B200_SEARCH_BLOCKS CFR_B200_DUST_SCREEN CFR_B200_L2_FETCH
