#!/bin/bash
# per-kernel numbers (ncu, few metrics) for environment-selected kernel variants
# usage: gpu_var.sh "<workloads>" "<variants>"   variant = comma-separated KEY=VALUE list (or "default")
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
M="gpu__time_duration.sum,smsp__inst_executed.sum,dram__bytes_read.sum,lts__t_sectors_srcunit_tex_op_read.sum,lts__t_requests_srcunit_tex_op_read.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__thread_inst_executed_per_inst_executed.ratio"
for W in $WL; do
  if [[ $W == m700* ]] && [ -n "$BUILD_PID" ]; then wait $BUILD_PID; BUILD_PID=""; fi
  for V in $VARS; do
    ENVS=""
    if [ "$V" != "default" ]; then ENVS=$(echo $V | tr ',' ' '); fi
    TAG=$(echo $V | tr -c 'A-Za-z0-9\n' '_')
    env $ENVS timeout 600 ncu --metrics $M --clock-control none -k regex:"k_search|k_locate|k_dust" -s 4 -c 4 --csv --log-file gpurun_out/var_${W}_${TAG}.csv python bench.py --workload $W --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/var_${W}_${TAG}.log 2>&1
    python - <<PY
import csv
rows=[r for r in csv.reader(open("gpurun_out/var_${W}_${TAG}.csv")) if len(r)>10 and r[0].isdigit()]
k={}
for r in rows:
    name=r[4].split("(")[0].replace("cfrb200::","").replace("void ","")
    k.setdefault((r[0],name),{})[r[-3]]=float(r[-1].replace(",",""))
out=[]
for (i,name),m in k.items():
    out.append("%s %.3fms inst %.0fM thr %.1f issue %.0f%% dram %.2fGB l2sec/req %.2f"%(name[:28], m.get("gpu__time_duration.sum",0)/1e6, m.get("smsp__inst_executed.sum",0)/1e6, m.get("smsp__thread_inst_executed_per_inst_executed.ratio",0), m.get("smsp__issue_active.avg.pct_of_peak_sustained_active",0), m.get("dram__bytes_read.sum",0)/1e9, m.get("lts__t_sectors_srcunit_tex_op_read.sum",0)/max(1,m.get("lts__t_requests_srcunit_tex_op_read.sum",1))))
print("${W} ${V}:"); print("   "+"\n   ".join(out))
PY
  done
done
