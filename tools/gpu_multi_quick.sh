#!/bin/bash
# two GPUs, quick: the multi-GPU tests and the CLI with --gpus 1 / 2 on 4 M pairs (same TSV)
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q ) > gpurun_out/pytest_gpu_multi.log 2>&1
tail -3 gpurun_out/pytest_gpu_multi.log
( time timeout 900 python tools/cli_scale.py c4 4 1 2 ) > gpurun_out/r02_cli_scale_c4.json 2> gpurun_out/r02_cli_scale_c4.err
tail -2 gpurun_out/r02_cli_scale_c4.err; cut -c1-1500 gpurun_out/r02_cli_scale_c4.json
