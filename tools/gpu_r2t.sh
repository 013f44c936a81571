#!/bin/bash
# the round's last GPU seconds: the by-reference witness mode — its GPU test, then the bench line with the labelled leg
mkdir -p gpurun_out
timeout 60 python -m pytest tests/test_zz_witness_by_reference.py -m gpu -x -q 2>&1 | tail -4
IPCFP_BENCH_NO_VERIFY=1 timeout 60 python bench.py --steps 10 --warmup 3 --no-storage > gpurun_out/r2t_bench_n1.json 2> gpurun_out/r2t_bench_n1.log
grep "resident timing\|by-reference" gpurun_out/r2t_bench_n1.log | cut -c1-200
