#!/bin/bash
# round 2: pair lines in k_search -- parity suite, then with / without on configs[2] (c3s = one batch) and configs[3]
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
    r=d["roofline"]
    print("$2: value %.4g e2e %.4g ms/step %.3f"%(d["value"], d["e2e"]["value"], d["ms_per_step"]), {k:round(v,3) for k,v in d.get("stage_ms_per_step",{}).items()}, "search useful GB/s %.0f"%r["achieved"], "pair bytes", d["details"].get("pair_line_bytes"), d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
except Exception as e:
    print(f, "FAILED", e); print(open(f.replace(".json",".err")).read()[-1500:])
PY
}
( time ensure c3 ) 2> gpurun_out/build_c3.time &
PID_C3=$!
( time timeout 1800 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
grep -E "passed|failed|rror" gpurun_out/pytest_gpu.log | tail -3
ensure c4
wait $PID_C3
for V in on off; do
  if [ $V = off ]; then export CFR_B200_PAIRS=0; else unset CFR_B200_PAIRS; fi
  timeout 900 python bench.py --workload c3s --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_c3s_pairs_$V.json 2> gpurun_out/r02_bench_c3s_pairs_$V.err
  show gpurun_out/r02_bench_c3s_pairs_$V.json "c3s pairs $V"
  timeout 900 python bench.py --workload c4 --reads 3000000 --steps 5 --warmup 2 --no-cpu-baseline > gpurun_out/r02_bench_c4_pairs_$V.json 2> gpurun_out/r02_bench_c4_pairs_$V.err
  show gpurun_out/r02_bench_c4_pairs_$V.json "c4 pairs $V"
done
unset CFR_B200_PAIRS
for SB in 10 12; do
  CFR_B200_PAIR_SEARCH_BLOCKS=$SB timeout 900 python bench.py --workload c3s --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_c3s_pairs_sb$SB.json 2> gpurun_out/r02_bench_c3s_pairs_sb$SB.err
  show gpurun_out/r02_bench_c3s_pairs_sb$SB.json "c3s pairs sb$SB"
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_search" -s 1 -c 1 -o gpurun_out/prof_r02_c3s_pairs -f python bench.py --workload c3s --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full_c3s_pairs.log 2>&1
tail -2 gpurun_out/ncu_full_c3s_pairs.log
timeout 900 python tests/cli_bench.py c3s 100000 > gpurun_out/r02_cli_c3_pairs_100000.json 2> gpurun_out/r02_cli_c3_pairs_100000.err
cat gpurun_out/r02_cli_c3_pairs_100000.json
