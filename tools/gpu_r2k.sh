#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_parallel.py -m gpu -x -q -k "nccl" 2>&1 | tail -5
IPCFP_XCH_TRACE=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 8 --warmup 3 --no-storage > gpurun_out/r2k_bench_n2.json 2> gpurun_out/r2k_bench_n2.log
grep "exchange:" gpurun_out/r2k_bench_n2.log | tail -4 | cut -c1-200
grep "resident timing\|parity" gpurun_out/r2k_bench_n2.log | cut -c1-200
