#!/bin/bash
# BASELINE configs[4] on ONE B200 -- 140 Gbp collection generated + indexed on the GPU, loaded, benched,
# and checked against the reference binary on a read sample
mkdir -p gpurun_out
export CFR_BUILD_VERBOSE=1
( time python -c "
import sys; sys.path.insert(0,'tools')
import make_data
print(make_data.ensure('c5'))" ) > gpurun_out/build_c5.log 2>&1
grep -E "cfr-build|real|built|rror|Traceback" gpurun_out/build_c5.log | tail -8
( time timeout 2400 python bench.py --workload c5 --steps 10 --warmup 3 --no-cpu-baseline ) > gpurun_out/r02_bench_c5.json 2> gpurun_out/r02_bench_c5.err
tail -4 gpurun_out/r02_bench_c5.err
cut -c1-3000 gpurun_out/r02_bench_c5.json
timeout 2400 python tests/cli_bench.py c5 20000 > gpurun_out/r02_cli_c5_20000.json 2> gpurun_out/r02_cli_c5_20000.err
cat gpurun_out/r02_cli_c5_20000.json; tail -3 gpurun_out/r02_cli_c5_20000.err
