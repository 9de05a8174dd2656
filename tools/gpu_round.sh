#!/bin/bash
# full measurement round: parity tests, smoke, benches (c2 / m700 / m700pe), launch list + full ncu capture
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
grep -E "passed|failed|error" gpurun_out/pytest_gpu.log | tail -3
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
for W in c2 m700 m700pe; do
  EXTRA="--no-cpu-baseline"; if [ $W = c2 ]; then EXTRA=""; fi
  timeout 900 python bench.py --workload $W --steps 5 --warmup 3 $EXTRA > gpurun_out/bench_$W.json 2> gpurun_out/bench_$W.err
  python - <<PY
import json
f="gpurun_out/bench_$W.json"
try:
    d=json.loads(open(f).read().strip().splitlines()[-1])
    print("$W: value %.4g e2e %.4g ms/step %.3f"%(d["value"], d["e2e"]["value"], d["ms_per_step"]), {k:round(v,3) for k,v in d["stage_ms_per_step"].items()}, "search alg GB/s %.0f"%d["roofline"]["achieved"], d.get("cpu_baseline"), d["clocks"])
except Exception as e:
    print(f, "FAILED", e); print(open(f.replace(".json",".err")).read()[-1500:])
PY
done
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_m700.csv python bench.py --workload m700 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launch_m700.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_c2.csv python bench.py --workload c2 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launch_c2.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_search|k_locate|k_dust|k_score|k_select|k_encode" -s 6 -c 6 -o gpurun_out/prof_r01_m700 -f python bench.py --workload m700 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu4.log 2>&1
tail -1 gpurun_out/ncu4.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_search|k_locate|k_dust|k_score|k_select|k_encode" -s 6 -c 6 -o gpurun_out/prof_r01_c2 -f python bench.py --workload c2 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_c2_full.log 2>&1
tail -1 gpurun_out/ncu_c2_full.log
