#!/bin/bash
mkdir -p gpurun_out
IPCFP_XCH_TRACE=1 IPCFP_BENCH_NO_VERIFY=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 6 --warmup 3 --no-storage > gpurun_out/r2i_bench_n2.json 2> gpurun_out/r2i_bench_n2.log
grep "exchange:" gpurun_out/r2i_bench_n2.log | tail -8
NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT,P2P IPCFP_BENCH_NO_VERIFY=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 2 --warmup 3 --no-storage > /dev/null 2> gpurun_out/r2i_nccl_info.log
grep -i "p2p\|nvls\|channel\|via" gpurun_out/r2i_nccl_info.log | head -30
