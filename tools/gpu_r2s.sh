#!/bin/bash
# last call of round 2 (budget ~2.8 GPU-min): bench.py at N=2 after the sampler moved in front of the warm-up (first timed step no longer cold)
mkdir -p gpurun_out
IPCFP_BENCH_NO_VERIFY=1 timeout 100 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 10 --warmup 3 --no-storage > gpurun_out/r2s_bench_n2.json 2> gpurun_out/r2s_bench_n2.log
grep "resident timing" gpurun_out/r2s_bench_n2.log | cut -c1-250 | tail -2
