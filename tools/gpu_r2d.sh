#!/bin/bash
# round 2, GPU call D (2 GPUs): bench at N=1 and N=2 (in-library NCCL protocol), with the parity verdicts
mkdir -p gpurun_out
timeout 900 python bench.py --gpus 1 --steps 10 --warmup 3 > gpurun_out/r2d_bench_n1.json 2> gpurun_out/r2d_bench_n1.log
tail -5 gpurun_out/r2d_bench_n1.log; cat gpurun_out/r2d_bench_n1.json | head -c 3000
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2d_bench_n2.json 2> gpurun_out/r2d_bench_n2.log
tail -12 gpurun_out/r2d_bench_n2.log; cat gpurun_out/r2d_bench_n2.json | head -c 3000
