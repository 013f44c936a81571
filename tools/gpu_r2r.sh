#!/bin/bash
# final 1-GPU call of round 2: GPU suite, ncu launch list of the bench command, clean N=1 bench line
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r2r_launches.csv python bench.py --steps 2 --warmup 3 --no-storage > gpurun_out/r2r_bench_under_ncu.log 2>&1
timeout 400 python bench.py --steps 10 --warmup 3 > gpurun_out/r2r_bench_n1.json 2> gpurun_out/r2r_bench_n1.log
grep "resident timing\|e2e timing" gpurun_out/r2r_bench_n1.log | cut -c1-260
