#!/bin/bash
# round 2, GPU call G: lean staged pass-1 variants (sweep + parity suite + ncu), verifier tests
mkdir -p gpurun_out
SWEEP="MINB=8,STAGE=lean128x4x1,STAGE=lean128x4x1w2,STAGE=lean128x4x2,STAGE=lean64x8x2,STAGE=128x4x1,MINB=8"
timeout 600 python tools/profile_step.py --steps 5 --warmup 2 --sweep "$SWEEP" > gpurun_out/r2g_sweep.txt 2>&1
grep SWEEP gpurun_out/r2g_sweep.txt
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2g_pytest_default.txt 2>&1
tail -15 gpurun_out/r2g_pytest_default.txt
IPCFP_PASS1_STAGE=lean128x4x1 timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2g_pytest_lean128x4x1.txt 2>&1
tail -3 gpurun_out/r2g_pytest_lean128x4x1.txt
IPCFP_PASS1_STAGE=lean128x4x1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_pass1 -s 2 -c 1 -f -o gpurun_out/r2g_ncu_lean128x4x1 python tools/profile_step.py --steps 1 --warmup 2 > gpurun_out/r2g_ncu.log 2>&1
