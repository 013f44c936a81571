python bench.py > gpurun_out/bench_r1_n1_v3.json 2> gpurun_out/bench_r1_n1_v3.err; cat gpurun_out/bench_r1_n1_v3.json; grep "e2e timing" gpurun_out/bench_r1_n1_v3.err | cut -c1-400
timeout 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "not full_size" -x 2>&1 | tail -2
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r1_launches_v4_bench.csv python bench.py --steps 2 --warmup 1 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on --kernel-name regex:k_pass1 -c 1 -s 3 -o gpurun_out/prof_r1_v4_pass1 python tools/profile_step.py --steps 1 --warmup 3 > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r1_launches_v4_step.csv python tools/profile_step.py --steps 2 --warmup 2 > /dev/null 2>&1
ls -la gpurun_out | tail -5
