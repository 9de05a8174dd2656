#!/bin/bash
# round 2: pair lines, one lane per search + warp-cooperative line fetch
mkdir -p gpurun_out
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
( time timeout 1800 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "variants or golden or small_vs_oracle" ) > gpurun_out/pytest_gpu_pairs.log 2>&1
grep -E "passed|failed|rror" gpurun_out/pytest_gpu_pairs.log | tail -3
for W in s2g c4; do
  R=""; if [ $W = c4 ]; then R="--reads 3000000 --steps 5"; else R="--steps 20"; fi
  for V in on off sb10; do
    unset CFR_B200_PAIRS CFR_B200_PAIR_SEARCH_BLOCKS
    if [ $V = off ]; then export CFR_B200_PAIRS=0; fi
    if [ $V = sb10 ]; then export CFR_B200_PAIR_SEARCH_BLOCKS=10; fi
    timeout 900 python bench.py --workload $W $R --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_${W}_pairs_$V.json 2> gpurun_out/r02_bench_${W}_pairs_$V.err
    show gpurun_out/r02_bench_${W}_pairs_$V.json "$W pairs $V"
  done
done
unset CFR_B200_PAIRS CFR_B200_PAIR_SEARCH_BLOCKS
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_search" -s 1 -c 1 -o gpurun_out/prof_r02_s2g_pairs -f python bench.py --workload s2g --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full_s2g_pairs.log 2>&1
tail -2 gpurun_out/ncu_full_s2g_pairs.log
timeout 900 python tests/cli_bench.py s2g 100000 > gpurun_out/r02_cli_s2g_pairs_100000.json 2> gpurun_out/r02_cli_s2g_pairs_100000.err
cat gpurun_out/r02_cli_s2g_pairs_100000.json
