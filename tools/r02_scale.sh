#!/bin/bash
# round 2: the driver's scaling run, N = 1, 2, 4, 8 back to back on one box (ours), reference arm at N = 8 once
mkdir -p gpurun_out
NG=$(nvidia-smi -L | wc -l)
for N in 1 2 4 8; do
  [ $N -gt $NG ] && continue
  if [ $N -eq 1 ]; then
    timeout 900 python bench.py --gpus 1 --steps 50 --warmup 5 > gpurun_out/r02g_bench_n$N.json 2> gpurun_out/r02g_bench_n$N.err
  else
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$N --master-addr 127.0.0.1 --master-port 2957$N bench.py --gpus $N --steps 50 --warmup 5 > gpurun_out/r02g_bench_n$N.json 2> gpurun_out/r02g_bench_n$N.err
  fi
  echo "N=$N rc=$?"
  python - $N <<'PY'
import json, sys
n = sys.argv[1]
try:
    d = json.loads(open(f"gpurun_out/r02g_bench_n{n}.json").read().strip().splitlines()[-1])
    print("  value %.4g ms_per_step %.4f fill_ms %.4f trace_ms %.4f roofline.frac %.3f e2e_ms %.4f launches %d parity %s" % (
        d["value"], d["ms_per_step"], d["fill_ms"], d["trace_ms"], d["roofline"]["frac"], d["e2e"]["ms_per_step"], d["gpu_launches"], d.get("parity_check")))
    if "parity_detail" in d: print("  ", d["parity_detail"])
    if "c4_trace_2160p" in d: print("  ", d["c4_trace_2160p"])
except Exception as e:
    print("no bench line:", e); print(open(f"gpurun_out/r02g_bench_n{n}.err").read()[-2500:])
PY
done
