#!/bin/bash
# round 2, N = 2: presenter's early row copies -- multi-GPU checks, the C++ host driver, bench with e2e
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_sharded_gpu.py tests/test_linked_gpu.py tests/test_cpp_host.py -x -q -m gpu > gpurun_out/r02q_tests.log 2>&1
echo "tests rc=$?"; tail -3 gpurun_out/r02q_tests.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29582 bench.py --gpus 2 --steps 30 --warmup 5 --no-extras > gpurun_out/r02q_bench_n2.json 2> gpurun_out/r02q_bench_n2.err
echo "bench rc=$?"; python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02q_bench_n2.json").read().strip().splitlines()[-1])
print("value %.4g ms_per_step %.4f fill_ms %.4f trace_ms %.4f frac %.3f e2e_ms %.4f parity %s" % (d["value"], d["ms_per_step"], d["fill_ms"], d["trace_ms"], d["roofline"]["frac"], d["e2e"]["ms_per_step"], d.get("parity_check")))
PY
tail -3 gpurun_out/r02q_bench_n2.err
