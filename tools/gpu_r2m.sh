#!/bin/bash
# 4-GPU box: NCCL test at world 2 and 4 (skewed-shard case included), N=4 bench with the protocol timeline
mkdir -p gpurun_out
timeout 600 python -m pytest "tests/test_parallel.py::test_sharded_call_over_nccl[2]" "tests/test_parallel.py::test_sharded_call_over_nccl[4]" -m gpu -x -q 2>&1 | tail -5
n=4
IPCFP_XCH_TRACE=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 8 --warmup 3 --no-storage > gpurun_out/r2m_bench_n$n.json 2> gpurun_out/r2m_bench_n$n.log
grep "timeline" gpurun_out/r2m_bench_n$n.log | tail -8 | cut -c1-260
grep "resident timing\|parity" gpurun_out/r2m_bench_n$n.log | cut -c1-220 | tail -3
