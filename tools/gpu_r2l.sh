#!/bin/bash
# 8-GPU box: NCCL tests at world 4 and 8, then the scaling points N=8 and N=4 (N=1/2 measured separately)
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 600 python -m pytest tests/test_parallel.py -m gpu -x -q -k "nccl" 2>&1 | tail -5
for n in 8 4; do
IPCFP_XCH_TRACE=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 8 --warmup 3 --no-storage > gpurun_out/r2l_bench_n$n.json 2> gpurun_out/r2l_bench_n$n.log
grep "rank 0\] exchange:" gpurun_out/r2l_bench_n$n.log | tail -2 | cut -c1-200
grep "resident timing\|parity" gpurun_out/r2l_bench_n$n.log | cut -c1-220 | tail -3
done
