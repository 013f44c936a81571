#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/profile_sharded.py --steps 4 > gpurun_out/r2f_sharded_w1.txt 2>&1
grep SHARDED gpurun_out/r2f_sharded_w1.txt
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2f_launches_sharded_w1.csv python tools/profile_sharded.py --steps 2 > gpurun_out/r2f_ncu.log 2>&1
python tools/ncu_summary.py gpurun_out/r2f_launches_sharded_w1.csv | head -60
