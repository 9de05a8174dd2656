#!/bin/bash
# short check of the drop-in CLI on one B200: parity tests, then FASTQ -> TSV wall time and byte
# equality with the reference binary on larger read files (m700 is built on the box meanwhile)
mkdir -p gpurun_out
( python -c "
import sys; sys.path.insert(0,'tools')
import make_data
make_data.ensure('m700')" > gpurun_out/build_m700.log 2>&1 ) &
PID=$!
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
grep -E "passed|failed|rror" gpurun_out/pytest_gpu.log | tail -2
wait $PID
for spec in "c2 ${1:-1000000}" "m700pe ${2:-500000}"; do
  F=gpurun_out/cli_$(echo $spec | tr ' ' '_')
  timeout 900 python tests/cli_bench.py $spec > $F.json 2> $F.err
  cat $F.json
done
