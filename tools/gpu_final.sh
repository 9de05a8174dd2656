#!/bin/bash
# closing check of a round: parity tests, smoke, the default bench invocation (both arms), HBM-resident benches
mkdir -p gpurun_out
( python -c "
import sys; sys.path.insert(0,'tools')
import make_data
make_data.ensure('m700')" > gpurun_out/build_m700.log 2>&1 ) &
PID=$!
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
grep -E "passed|failed|rror" gpurun_out/pytest_gpu.log | tail -2
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
wait $PID
timeout 900 python bench.py > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_c2_reference.json 2> gpurun_out/bench_c2_reference.err
for W in m700 m700pe; do
  timeout 900 python bench.py --workload $W --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$W.json 2> gpurun_out/bench_$W.err
done
python - <<'PY'
import json
for w in ("c2", "c2_reference", "m700", "m700pe"):
    try:
        d = json.loads(open("gpurun_out/bench_%s.json" % w).read().strip().splitlines()[-1])
        print(w, "value %.4g" % d["value"], "e2e %.4g" % d["e2e"]["value"], "ms/step %.3f" % d["ms_per_step"], d.get("cpu_baseline", {}).get("value"))
    except Exception as e:
        print(w, "FAILED", e)
PY
