#!/bin/bash
# round 2, GPU call B: in-library NCCL protocol at world 1 (every kernel + NCCL call of the sharded path on one GPU), full GPU suite
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_parallel.py -m gpu -x -q -k "nccl" > gpurun_out/r2b_pytest_nccl.txt 2>&1
tail -40 gpurun_out/r2b_pytest_nccl.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2b_pytest_all.txt 2>&1
tail -5 gpurun_out/r2b_pytest_all.txt
