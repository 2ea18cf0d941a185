#!/bin/bash
# round 2, GPU call 3 (1 GPU): whole GPU suite, smoke, N=1 bench line with the C3/C5 extras
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r02c_gpu_tests.log 2>&1
echo "gpu tests rc=$?"; tail -5 gpurun_out/r02c_gpu_tests.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/r02c_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r02c_smoke.log
timeout 900 python bench.py > gpurun_out/r02c_bench_n1.json 2> gpurun_out/r02c_bench_n1.err; echo "bench rc=$?"
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r02c_bench_n1.json").read().strip().splitlines()[-1])
    for k in ("value", "ms_per_step", "fill_ms", "trace_ms", "roofline", "e2e", "csg_1k_512", "dirty_60hz", "trace_closeup", "trace_modes", "mesh"):
        print(k, json.dumps(d.get(k)))
except Exception as e:
    print("no bench line:", e); print(open("gpurun_out/r02c_bench_n1.err").read()[-3000:])
PY
