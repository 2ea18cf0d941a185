#!/bin/bash
# round 2, GPU call 4 (2 GPUs): N=2 bench through linked slabs, the round-1 path beside it, reference arm at N=2
mkdir -p gpurun_out
run() { # name, extra args
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 2 --steps 50 --warmup 5 $2 > gpurun_out/$1.json 2> gpurun_out/$1.err
  echo "$1 rc=$?"
  python - "$1" <<'PY'
import json, sys
n = sys.argv[1]
try:
    d = json.loads(open(f"gpurun_out/{n}.json").read().strip().splitlines()[-1])
    for k in ("value", "ms_per_step", "fill_ms", "trace_ms", "roofline", "e2e", "parity_check", "parity_detail", "c4_trace_2160p", "gpu_launches", "detail"):
        if k in d: print(" ", k, json.dumps(d.get(k)))
except Exception as e:
    print("no bench line:", e); print(open(f"gpurun_out/{n}.err").read()[-3000:])
PY
}
run r02d_bench_n2_linked ""
run r02d_bench_n2_round1 "--no-linked --no-extras"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29552 bench.py --gpus 2 --steps 2 --warmup 1 --impl reference > gpurun_out/r02d_ref_n2.json 2> gpurun_out/r02d_ref_n2.err; echo "ref rc=$?"; cut -c1-600 gpurun_out/r02d_ref_n2.json
