#!/bin/bash
# 8-GPU box: NCCL test at world 8, scaling points N=8 and N=4
mkdir -p gpurun_out
timeout 400 python -m pytest "tests/test_parallel.py::test_sharded_call_over_nccl[8]" -m gpu -x -q 2>&1 | tail -4
for n in 8 4; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 10 --warmup 3 --no-storage > gpurun_out/r2q_bench_n$n.json 2> gpurun_out/r2q_bench_n$n.log
grep "resident timing\|parity\|e2e timing" gpurun_out/r2q_bench_n$n.log | cut -c1-250 | tail -3
done
