#!/bin/bash
# quick GPU iteration: parity tests then the bench in the variants named by $VARIANTS
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
tail -6 gpurun_out/pytest_gpu.log
W=${WORKLOAD:-c2}
timeout 600 python bench.py --workload $W --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${W}_coop.json 2> gpurun_out/bench_${W}_coop.err
python - <<PY
import json
for f in ["gpurun_out/bench_${W}_coop.json"]:
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "value %.3g e2e %.3g ms/step %.3f"%(d["value"], d["e2e"]["value"], d["ms_per_step"]), d["stage_ms_per_step"], "roofline", d["roofline"]["achieved"], d["clocks"])
    except Exception as e:
        print(f, "FAILED", e); print(open(f.replace(".json",".err")).read()[-2000:])
PY
CFR_B200_SCALAR_OCC=1 timeout 600 python bench.py --workload $W --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${W}_scalar.json 2> gpurun_out/bench_${W}_scalar.err
python - <<PY
import json
for f in ["gpurun_out/bench_${W}_scalar.json"]:
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "value %.3g e2e %.3g ms/step %.3f"%(d["value"], d["e2e"]["value"], d["ms_per_step"]), d["stage_ms_per_step"])
    except Exception as e:
        print(f, "FAILED", e); print(open(f.replace(".json",".err")).read()[-2000:])
PY
