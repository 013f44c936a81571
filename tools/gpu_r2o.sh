#!/bin/bash
# 2-GPU box: the whole GPU suite (rank-space witness bitmap, partitioned union), N=1 and N=2 bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2o_bench_n1.json 2> gpurun_out/r2o_bench_n1.log
grep "resident timing\|e2e timing\|storage" gpurun_out/r2o_bench_n1.log | cut -c1-260 | tail -4
n=2
IPCFP_XCH_TRACE=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 8 --warmup 3 --no-storage > gpurun_out/r2o_bench_n$n.json 2> gpurun_out/r2o_bench_n$n.log
grep "resident timing\|parity" gpurun_out/r2o_bench_n$n.log | cut -c1-220 | tail -3
