#!/bin/bash
# full measurement round on one B200: parity tests, smoke, ncu captures, benches (c2 / m700 / m700pe / c3).
# m700 and c3 are built on the box by the reference builder (cheaper than shipping them); the clean
# bench runs start only after the builds are done so that the host cores are idle.
mkdir -p gpurun_out
ensure() { python -c "
import sys; sys.path.insert(0,'tools')
import make_data
print(make_data.ensure('$1'))" > gpurun_out/build_$1.log 2>&1; }
show() {
python - <<PY
import json
f="$1"
try:
    d=json.loads(open(f).read().strip().splitlines()[-1])
    print("$2: value %.4g e2e %.4g ms/step %.3f"%(d["value"], d["e2e"]["value"], d["ms_per_step"]), {k:round(v,3) for k,v in d.get("stage_ms_per_step",{}).items()}, "search alg GB/s %.0f"%d["roofline"]["achieved"], d.get("cpu_baseline"), d["clocks"])
except Exception as e:
    print(f, "FAILED", e); print(open(f.replace(".json",".err")).read()[-1500:])
PY
}
( time ensure m700 ) 2> gpurun_out/build_m700.time &
PID_M700=$!
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
grep -E "passed|failed|rror" gpurun_out/pytest_gpu.log | tail -3
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
wait $PID_M700
if [ "$1" = "c3" ]; then ( time ensure c3 ) 2> gpurun_out/build_c3.time & PID_C3=$!; fi
# ---- ncu (GPU-side durations; host load from the c3 build does not matter here)
for W in c2 m700; do
  timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches_$W.csv python bench.py --workload $W --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launch_$W.log 2>&1
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_search|k_locate|k_dust|k_score|k_select|k_encode" -s 7 -c 7 -o gpurun_out/prof_r01_$W -f python bench.py --workload $W --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full_$W.log 2>&1
  tail -1 gpurun_out/ncu_full_$W.log
done
if [ -n "$PID_C3" ]; then wait $PID_C3; cat gpurun_out/build_c3.time | tail -3; fi
# ---- clean benches
timeout 900 python bench.py --workload c2 --steps 10 --warmup 3 > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err
show gpurun_out/bench_c2.json c2
CFR_B200_TRACE=1 timeout 600 python bench.py --workload c2 --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c2_trace.json 2> gpurun_out/bench_c2_trace.err
grep "cfr trace" gpurun_out/bench_c2_trace.err | head -12
timeout 600 python bench.py --impl reference --workload c2 --steps 2 --warmup 1 > gpurun_out/bench_c2_reference.json 2> gpurun_out/bench_c2_reference.err
cat gpurun_out/bench_c2_reference.json | cut -c1-400
for W in m700 m700pe; do
  timeout 900 python bench.py --workload $W --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$W.json 2> gpurun_out/bench_$W.err
  show gpurun_out/bench_$W.json $W
done
if [ "$1" = "c3" ]; then
  timeout 1200 python bench.py --workload c3 --steps 5 --warmup 3 > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err
  show gpurun_out/bench_c3.json c3
  ( time timeout 900 python -m pytest tests/test_gpu_properties.py -m gpu -x -q -k c3 ) > gpurun_out/pytest_c3.log 2>&1
  tail -3 gpurun_out/pytest_c3.log
fi
# ---- drop-in CLI: FASTQ -> TSV wall time and byte equality with the reference binary at scale
for spec in "c2 500000" "m700pe 100000"; do
  timeout 900 python tests/cli_bench.py $spec > gpurun_out/cli_$(echo $spec | tr ' ' '_').json 2> gpurun_out/cli_$(echo $spec | tr ' ' '_').err
  cat gpurun_out/cli_$(echo $spec | tr ' ' '_').json
done
if [ "$1" = "c3" ]; then
  timeout 1200 python tests/cli_bench.py c3 100000 > gpurun_out/cli_c3_100000.json 2> gpurun_out/cli_c3_100000.err
  cat gpurun_out/cli_c3_100000.json
fi
