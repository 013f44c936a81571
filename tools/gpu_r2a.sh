#!/bin/bash
# round 2, GPU call A: pass-1 variant sweep (the design decision), GPU parity suite on the default and on the staged kernel, ncu captures
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2a_env.txt 2>&1
SWEEP="MINB=8,STAGE=128x4x1,STAGE=128x4x1w2,STAGE=128x4x2,STAGE=64x8x2,STAGE=64x4x2,STAGE=256x2x1,STAGE=256x4x1,RING=128x2,RING=128x4,RING=256x2,MINB=8"
timeout 600 python tools/profile_step.py --steps 5 --warmup 2 --sweep "$SWEEP" > gpurun_out/r2a_sweep.txt 2>&1
grep SWEEP gpurun_out/r2a_sweep.txt
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2a_pytest_default.txt 2>&1
tail -3 gpurun_out/r2a_pytest_default.txt
IPCFP_PASS1_STAGE=128x4x1 timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2a_pytest_stage128x4x1.txt 2>&1
tail -3 gpurun_out/r2a_pytest_stage128x4x1.txt
IPCFP_PASS1_STAGE=64x8x2 timeout 900 python -m pytest tests -m gpu -x -q -k "event or error or shape" > gpurun_out/r2a_pytest_stage64x8x2.txt 2>&1
tail -3 gpurun_out/r2a_pytest_stage64x8x2.txt
for v in 128x4x1 64x8x2; do
  IPCFP_PASS1_STAGE=$v timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_pass1 -s 2 -c 1 -f -o gpurun_out/r2a_ncu_stage_$v python tools/profile_step.py --steps 1 --warmup 2 > gpurun_out/r2a_ncu_stage_$v.log 2>&1
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_pass1 -s 2 -c 1 -f -o gpurun_out/r2a_ncu_occ8 python tools/profile_step.py --steps 1 --warmup 2 > gpurun_out/r2a_ncu_occ8.log 2>&1
ls -la gpurun_out | tail -20
