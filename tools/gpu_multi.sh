#!/bin/bash
# two GPUs: the multi-GPU tests, the bench under torchrun at N = 2 (the driver's launch line), the CLI with --gpus 1 / 2
mkdir -p gpurun_out
nvidia-smi -L
( time timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q ) > gpurun_out/pytest_gpu_multi.log 2>&1
tail -3 gpurun_out/pytest_gpu_multi.log
( time timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 20 --warmup 3 ) > gpurun_out/r02_bench_c4_2gpu.json 2> gpurun_out/r02_bench_c4_2gpu.err
tail -3 gpurun_out/r02_bench_c4_2gpu.err; cut -c1-1200 gpurun_out/r02_bench_c4_2gpu.json
( time timeout 1500 python tools/cli_scale.py c4 4 1 2 ) > gpurun_out/r02_cli_scale_c4.json 2> gpurun_out/r02_cli_scale_c4.err
tail -3 gpurun_out/r02_cli_scale_c4.err; cut -c1-3000 gpurun_out/r02_cli_scale_c4.json
