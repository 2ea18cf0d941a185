#!/bin/bash
# round 2: the GPU suite and smoke of the final build (one GPU)
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -x -q -m gpu > gpurun_out/r02y_gpu_tests.log 2>&1; echo "gpu tests rc=$?"; tail -3 gpurun_out/r02y_gpu_tests.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02y_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r02y_smoke.log
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r02y_bench_n1.json 2> gpurun_out/r02y_bench_n1.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02y_bench_n1.json").read().strip().splitlines()[-1])
print("value %.4g ms_per_step %.4f fill_ms %.4f trace_ms %.4f frac %.3f e2e_ms %.4f traffic %s" % (d["value"], d["ms_per_step"], d["fill_ms"], d["trace_ms"], d["roofline"]["frac"], d["e2e"]["ms_per_step"], d["roofline"]["traffic"]))
print(d["roofline"]["traffic_source"]); print(d["trace_profile"])
PY
