#!/bin/bash
# round 2, N = 1: tile order from the previous frame -- trace tests, timing on / off
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_trace_gpu.py tests/test_full_size_gpu.py tests/test_abi.py -x -q -m gpu > gpurun_out/r02r_trace_tests.log 2>&1
echo "tests rc=$?"; tail -3 gpurun_out/r02r_trace_tests.log
python - <<'PY'
import sys, time, torch
sys.path.insert(0, ".")
import sdf_viewer_b200 as S
BB = ((-1.0, -1.0, -1.0), (1.0, 1.0, 1.0))
W, H = 1920, 1080
with S.SDFViewer.new_voxels((512, 512, 512), BB, 2) as v:
    v.set_tape(S.tape.demo_tape()); v.fill_all(); v.commit()
    stream = torch.cuda.ExternalStream(v.stream)
    r = torch.empty((H, W, 4), dtype=torch.uint8).pin_memory().numpy(); d = torch.empty((H, W), dtype=torch.float32).pin_memory().numpy()
    for name, cam in (("default", S.default_camera(W, H)), ("closeup", S.look_at_camera((0.9, 1.1, 1.8), (0, 0, 0), W, H)),
                      ("along -z", S.look_at_camera((0.1, 0.05, 4.0), (0, 0, 0), W, H))):
        out = []
        for order in (0, 1, 2):
            v.set_option("trace_tile_order", order)
            for _ in range(3): v.trace_device(cam, W, H)
            v.sync()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for _ in range(30): v.trace_device(cam, W, H)
            e1.record(stream); v.sync(); torch.cuda.synchronize()
            for _ in range(3): v.trace_rgba8(cam, W, H, r, d)
            t = time.perf_counter()
            for _ in range(30): v.trace_rgba8(cam, W, H, r, d)
            out.append(f"order {order}: device {e0.elapsed_time(e1) / 30:.4f} ms, trace+D2H {(time.perf_counter() - t) / 30 * 1e3:.4f} ms")
        print(f"{name}: " + "   ".join(out), flush=True)
PY
timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-extras > gpurun_out/r02r_bench_n1.json 2> gpurun_out/r02r_bench_n1.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02r_bench_n1.json").read().strip().splitlines()[-1])
print("value %.4g ms_per_step %.4f fill_ms %.4f trace_ms %.4f frac %.3f e2e_ms %.4f launches %d" % (d["value"], d["ms_per_step"], d["fill_ms"], d["trace_ms"], d["roofline"]["frac"], d["e2e"]["ms_per_step"], d["gpu_launches"]))
PY
