#!/bin/bash
# round 2, N = 2: after the presenter fix -- multi-GPU checks, back-to-back 4K frames, a short bench with the C4 leg
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_sharded_gpu.py tests/test_linked_gpu.py -x -q -m gpu > gpurun_out/r02p_linked_tests.log 2>&1
echo "tests rc=$?"; tail -3 gpurun_out/r02p_linked_tests.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29582 bench.py --gpus 2 --steps 30 --warmup 5 > gpurun_out/r02p_bench_n2.json 2> gpurun_out/r02p_bench_n2.err
echo "bench rc=$?"; python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02p_bench_n2.json").read().strip().splitlines()[-1])
print("value %.4g ms_per_step %.4f fill_ms %.4f trace_ms %.4f frac %.3f e2e_ms %.4f parity %s" % (d["value"], d["ms_per_step"], d["fill_ms"], d["trace_ms"], d["roofline"]["frac"], d["e2e"]["ms_per_step"], d.get("parity_check")))
print(d.get("c4_trace_2160p")); print(d.get("alternatives"))
PY
tail -3 gpurun_out/r02p_bench_n2.err
