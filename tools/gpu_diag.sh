#!/bin/bash
# diagnosis round: parity tests, c2 bench + launch list + full ncu capture, m700 built on the box in the
# background (cheaper than shipping it), then DRAM-granularity metrics of k_search on m700
mkdir -p gpurun_out
nproc > gpurun_out/nproc.txt
( time python -c "
import sys; sys.path.insert(0,'tools')
import make_data
print(make_data.ensure('m700'))
" ) > gpurun_out/m700_build.log 2>&1 &
BUILD_PID=$!
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
grep -E "passed|failed|rror" gpurun_out/pytest_gpu.log | tail -3
show() {
python - <<PY
import json
f="$1"
try:
    d=json.loads(open(f).read().strip().splitlines()[-1])
    print("$2: value %.4g e2e %.4g ms/step %.3f"%(d["value"], d["e2e"]["value"], d["ms_per_step"]), {k:round(v,3) for k,v in d["stage_ms_per_step"].items()}, "search GB/s %.0f"%d["roofline"]["achieved"], d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
except Exception as e:
    print(f, "FAILED", e); print(open(f.replace(".json",".err")).read()[-1500:])
PY
}
timeout 600 python bench.py --workload c2 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err
show gpurun_out/bench_c2.json c2
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches_c2.csv python bench.py --workload c2 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launch_c2.log 2>&1
grep -E "k_|Kernel" gpurun_out/launches_c2.csv | grep "gpu__time_duration" | cut -d, -f5,12- | head -24
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_search|k_dust|k_locate" -s 4 -c 4 -o gpurun_out/prof_c2 -f python bench.py --workload c2 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full_c2.log 2>&1
tail -1 gpurun_out/ncu_full_c2.log
wait $BUILD_PID
tail -3 gpurun_out/m700_build.log
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__sectors_read.sum,lts__t_sectors_srcunit_tex_op_read.sum,lts__t_sectors_srcunit_tex_op_read_lookup_miss.sum,lts__t_requests_srcunit_tex_op_read.sum,lts__t_sectors_srcunit_tex_op_read_lookup_hit.sum"
for G in 0 32 128; do
  CFR_B200_L2_FETCH=$G timeout 600 ncu --metrics $M --clock-control none -k regex:"k_search|k_locate" -s 2 -c 2 --csv --log-file gpurun_out/m700_l2_$G.csv python bench.py --workload m700 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_m700_$G.log 2>&1
  echo "== L2 fetch $G"; grep -E "k_search|k_locate" gpurun_out/m700_l2_$G.csv | cut -d, -f5,13- | sed 's/"//g' | head -20
done
timeout 600 python bench.py --workload m700 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_m700.json 2> gpurun_out/bench_m700.err
show gpurun_out/bench_m700.json m700
