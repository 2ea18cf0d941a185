#!/bin/bash
# round 2, N = 8: the bench line of the default linked path (local halos, streaming trace) + link timing
mkdir -p gpurun_out
N=8
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$N --master-addr 127.0.0.1 --master-port 29578 bench.py --gpus $N --steps 50 --warmup 5 > gpurun_out/r02k_bench_n$N.json 2> gpurun_out/r02k_bench_n$N.err
echo "bench N=$N rc=$?"
python - $N <<'PY'
import json, sys
n = sys.argv[1]
try:
    d = json.loads(open(f"gpurun_out/r02k_bench_n{n}.json").read().strip().splitlines()[-1])
    print("  value %.4g ms_per_step %.4f fill_ms %.4f trace_ms %.4f roofline.frac %.3f e2e_ms %.4f launches %d parity %s" % (
        d["value"], d["ms_per_step"], d["fill_ms"], d["trace_ms"], d["roofline"]["frac"], d["e2e"]["ms_per_step"], d["gpu_launches"], d.get("parity_check")))
    for k in ("parity_detail", "c4_trace_2160p", "alternatives"):
        if k in d: print("  ", k, d[k])
except Exception as e:
    print("no bench line:", e); print(open(f"gpurun_out/r02k_bench_n{n}.err").read()[-2500:])
PY
SDFGPU_LINK_TIMING=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$N --master-addr 127.0.0.1 --master-port 29581 tools/link_timing.py 512 0 0 > gpurun_out/r02k_timing_n$N.log 2>&1
echo "timing rc=$?"; grep "^==" gpurun_out/r02k_timing_n$N.log
