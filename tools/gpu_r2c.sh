#!/bin/bash
# round 2, GPU call C (2 GPUs): in-library NCCL protocol at world 2
mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m pytest tests/test_parallel.py -m gpu -x -q -k "nccl" > gpurun_out/r2c_pytest_nccl_2gpu.txt 2>&1
tail -40 gpurun_out/r2c_pytest_nccl_2gpu.txt
