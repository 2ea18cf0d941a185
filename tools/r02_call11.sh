#!/bin/bash
# round 2, N = 1: trace tests (bands), bench line
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_trace_gpu.py tests/test_abi.py -x -q -m gpu > gpurun_out/r02l_trace_tests.log 2>&1
echo "tests rc=$?"; tail -3 gpurun_out/r02l_trace_tests.log
for b in 1 6; do
python - $b <<'PY'
import sys, time, numpy as np, torch
sys.path.insert(0, "."); 
import sdf_viewer_b200 as S
b = int(sys.argv[1])
BB = ((-1.0, -1.0, -1.0), (1.0, 1.0, 1.0))
W, H = 1920, 1080
with S.SDFViewer.new_voxels((512, 512, 512), BB, 2) as v:
    v.set_tape(S.tape.demo_tape()); v.fill_all(); v.commit()
    v.set_option("trace_bands", b)
    r = torch.empty((H, W, 4), dtype=torch.uint8).pin_memory().numpy(); d = torch.empty((H, W), dtype=torch.float32).pin_memory().numpy()
    for name, cam in (("default", S.default_camera(W, H)), ("closeup", S.look_at_camera((0.9, 1.1, 1.8), (0, 0, 0), W, H))):
        for _ in range(5): v.trace_rgba8(cam, W, H, r, d)
        t = time.perf_counter()
        n = 50
        for _ in range(n): v.trace_rgba8(cam, W, H, r, d)
        print(f"bands {b} {name}: trace + D2H {(time.perf_counter() - t) / n * 1e3:.3f} ms", flush=True)
PY
done
timeout 900 python bench.py --steps 50 --warmup 5 > gpurun_out/r02l_bench_n1.json 2> gpurun_out/r02l_bench_n1.err
echo "bench rc=$?"; python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02l_bench_n1.json").read().strip().splitlines()[-1])
print("value %.4g ms_per_step %.4f fill_ms %.4f trace_ms %.4f frac %.3f e2e_ms %.4f" % (d["value"], d["ms_per_step"], d["fill_ms"], d["trace_ms"], d["roofline"]["frac"], d["e2e"]["ms_per_step"]))
PY
tail -3 gpurun_out/r02l_bench_n1.err
