#!/bin/bash
mkdir -p gpurun_out
SWEEP="MINB=8,W16=1,W16=6,W16=1+TUNE=2,W16=1+TUNE=48,W16=1+TUNE=8,MINB=8"
timeout 600 python tools/profile_step.py --steps 5 --warmup 2 --sweep "$SWEEP" > gpurun_out/r2h_sweep.txt 2>&1
grep SWEEP gpurun_out/r2h_sweep.txt
IPCFP_PASS1_W16=1 timeout 900 python -m pytest tests -m gpu -x -q -k "event or error or shape or full" > gpurun_out/r2h_pytest_w16.txt 2>&1
tail -3 gpurun_out/r2h_pytest_w16.txt
IPCFP_PASS1_W16=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_pass1 -s 2 -c 1 -f -o gpurun_out/r2h_ncu_w16 python tools/profile_step.py --steps 1 --warmup 2 > gpurun_out/r2h_ncu.log 2>&1
