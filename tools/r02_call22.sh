#!/bin/bash
# round 2, N = 2: bench.py's fallback path (a rank "fails" in the warm-up -> every rank re-links with the trace in rounds)
mkdir -p gpurun_out
SDFGPU_BENCH_FORCE_FALLBACK=1 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29582 bench.py --gpus 2 --steps 10 --warmup 3 --no-extras > gpurun_out/r02u_bench_fallback_n2.json 2> gpurun_out/r02u_bench_fallback_n2.err
echo "bench rc=$?"; python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02u_bench_fallback_n2.json").read().strip().splitlines()[-1])
print("ms_per_step %.4f fill_ms %.4f trace_ms %.4f e2e_ms %.4f parity %s" % (d["ms_per_step"], d["fill_ms"], d["trace_ms"], d["e2e"]["ms_per_step"], d.get("parity_check")))
print(d.get("fallback")); print(d["detail"]["sharding"])
PY
tail -4 gpurun_out/r02u_bench_fallback_n2.err
