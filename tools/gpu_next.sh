#!/bin/bash
# The first GPU call a next session should make (one B200, ~12 min): what the last session of round 2 changed without a GPU.
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash tools/gpu_next.sh'
set -x
mkdir -p gpurun_out
# 1. the GPU suite on the default build (k_pass2 per warp, verify_items.cuh, the C++ host mirror, the new known-answer tests)
python -m pytest tests -m gpu -x -q > gpurun_out/next_gpu_tests.log 2>&1; tail -3 gpurun_out/next_gpu_tests.log
# 2. A/B of the k_pass2 launch shape: device_ms_breakdown.pass2 and ms_per_step, per warp (default) vs per thread (measured in round 2)
python bench.py --steps 10 --warmup 3 --no-storage > gpurun_out/next_bench_pass2_per_warp.json 2> gpurun_out/next_bench_a.log
IPCFP_PASS2_PER_THREAD=1 python bench.py --steps 10 --warmup 3 --no-storage > gpurun_out/next_bench_pass2_per_thread.json 2> gpurun_out/next_bench_b.log
python - <<'PY'
import json
for k in ("per_warp", "per_thread"):
    d = json.load(open(f"gpurun_out/next_bench_pass2_{k}.json"))
    print(k, "ms/step", round(d["ms_per_step"], 4), "pass2 phase", round(d["device_ms_breakdown"]["pass2"], 4), "parity", d["parity"])
PY
# 3. launch list of the bench command (k_pass2's own duration in both shapes)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/next_launches.csv python bench.py --steps 2 --warmup 1 --no-storage --no-cpu-baseline > /dev/null 2>&1
grep -i "k_pass2" gpurun_out/next_launches.csv | tail -3
# 4. the library with only the C ABI exported (make HIDE_INTERNALS=1): same suite, then it can become the default
make HIDE_INTERNALS=1 LIB_OUT=/tmp/libipcfp_hidden.so /tmp/libipcfp_hidden.so > /dev/null && cp ipc_filecoin_proofs_b200/libipcfp.so /tmp/libipcfp_default.so && cp /tmp/libipcfp_hidden.so ipc_filecoin_proofs_b200/libipcfp.so
python -m pytest tests -m gpu -x -q > gpurun_out/next_gpu_tests_hidden.log 2>&1; tail -3 gpurun_out/next_gpu_tests_hidden.log
cp /tmp/libipcfp_default.so ipc_filecoin_proofs_b200/libipcfp.so
