#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L | wc -l; free -g | head -2; nproc
( time timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 4 --steps 10 --warmup 3 ) > gpurun_out/r02_bench_c4_4gpu.json 2> gpurun_out/r02_bench_c4_4gpu.err
tail -6 gpurun_out/r02_bench_c4_4gpu.err; cut -c1-600 gpurun_out/r02_bench_c4_4gpu.json
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29520 bench.py --impl reference --gpus 4 --steps 3 --warmup 1 ) > gpurun_out/r02_bench_c4_4gpu_ref.json 2> gpurun_out/r02_bench_c4_4gpu_ref.err
tail -2 gpurun_out/r02_bench_c4_4gpu_ref.err; cut -c1-300 gpurun_out/r02_bench_c4_4gpu_ref.json
