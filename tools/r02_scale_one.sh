#!/bin/bash
# round 2: one line of the scaling run (the driver's command) on a box with exactly N GPUs
N=$1
mkdir -p gpurun_out
if [ $N -eq 1 ]; then
  timeout 900 python bench.py --gpus 1 --steps 50 --warmup 5 > gpurun_out/r02s_bench_n$N.json 2> gpurun_out/r02s_bench_n$N.err
  echo "N=$N rc=$?"
  timeout 600 python bench.py --impl reference --gpus 1 --steps 2 --warmup 1 > gpurun_out/r02s_reference_n1.json 2> gpurun_out/r02s_reference_n1.err
  echo "reference rc=$?"; cut -c1-400 gpurun_out/r02s_reference_n1.json
else
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$N --master-addr 127.0.0.1 --master-port 2957$N bench.py --gpus $N --steps 50 --warmup 5 > gpurun_out/r02s_bench_n$N.json 2> gpurun_out/r02s_bench_n$N.err
  echo "N=$N rc=$?"
fi
python - $N <<'PY'
import json, sys
n = sys.argv[1]
try:
    d = json.loads(open(f"gpurun_out/r02s_bench_n{n}.json").read().strip().splitlines()[-1])
    print("  value %.4g ms_per_step %.4f fill_ms %.4f trace_ms %.4f roofline.frac %.3f e2e_ms %.4f launches %d parity %s" % (
        d["value"], d["ms_per_step"], d["fill_ms"], d["trace_ms"], d["roofline"]["frac"], d["e2e"]["ms_per_step"], d["gpu_launches"], d.get("parity_check")))
    for k in ("parity_detail", "c4_trace_2160p", "alternatives", "csg_1k_512", "dirty_60hz", "mesh", "trace_closeup", "clocks"):
        if k in d: print("  ", k, json.dumps(d[k])[:600])
except Exception as e:
    print("no bench line:", e); print(open(f"gpurun_out/r02s_bench_n{n}.err").read()[-2500:])
PY
