#!/bin/bash
# 2-GPU box: NCCL tests (fixed-slot partitioned union incl. the forced overflow repeat), N=2 bench with the timeline
mkdir -p gpurun_out
timeout 600 python -m pytest "tests/test_parallel.py::test_sharded_call_over_nccl[1]" "tests/test_parallel.py::test_sharded_call_over_nccl[2]" -m gpu -x -q 2>&1 | tail -5
n=2
IPCFP_XCH_TRACE=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 8 --warmup 3 --no-storage > gpurun_out/r2p_bench_n$n.json 2> gpurun_out/r2p_bench_n$n.log
grep "resident timing\|parity\|e2e timing" gpurun_out/r2p_bench_n$n.log | cut -c1-250 | tail -4
grep "rank 0\] timeline" gpurun_out/r2p_bench_n$n.log | sed -n 5,8p | cut -c1-280
