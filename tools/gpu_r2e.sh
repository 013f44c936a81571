#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/profile_step.py --receipts 20000 --steps 1 --warmup 1 --storage 1000 > gpurun_out/r2e_storage.txt 2>&1
grep STORAGE gpurun_out/r2e_storage.txt | sed -n '2p;6p;10p'
timeout 600 python tools/profile_step.py --receipts 20000 --steps 1 --warmup 1 --storage 65536 > gpurun_out/r2e_storage64k.txt 2>&1
grep STORAGE gpurun_out/r2e_storage64k.txt | sed -n '2p;6p;10p'
