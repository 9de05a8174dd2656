#!/bin/bash
# round 2: the GPU index builder -- parity with the reference builder, then the 20 Gbp collection (configs[3])
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests/test_builder.py -m gpu -x -q ) > gpurun_out/pytest_builder.log 2>&1
tail -5 gpurun_out/pytest_builder.log
export CFR_BUILD_VERBOSE=1
for D in s400 ${1:-c4}; do
  ( time python -c "
import sys; sys.path.insert(0,'tools')
import make_data
print(make_data.ensure('$D'))" ) > gpurun_out/build_$D.log 2>&1
  grep -E "cfr-build|real|built|rror" gpurun_out/build_$D.log | tail -12
  ls -la data/$D/ | tail -6
done
nvidia-smi --query-gpu=memory.used --format=csv
W=${1:-c4}
timeout 1500 python bench.py --workload $W --steps 3 --warmup 1 --reads 3000000 --no-cpu-baseline > gpurun_out/r02_bench_${W}_first.json 2> gpurun_out/r02_bench_${W}_first.err
tail -5 gpurun_out/r02_bench_${W}_first.err
cut -c1-2500 gpurun_out/r02_bench_${W}_first.json
timeout 1500 python tests/cli_bench.py $W 50000 > gpurun_out/r02_cli_${W}_50000.json 2> gpurun_out/r02_cli_${W}_50000.err
cat gpurun_out/r02_cli_${W}_50000.json; tail -3 gpurun_out/r02_cli_${W}_50000.err
