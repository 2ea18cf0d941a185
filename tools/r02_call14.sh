#!/bin/bash
# round 2, N = 1: speculative next-cell fetch of long marches -- trace tests, timing against the step count it starts at
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_trace_gpu.py tests/test_full_size_gpu.py -x -q -m gpu > gpurun_out/r02o_trace_tests.log 2>&1
echo "tests rc=$?"; tail -3 gpurun_out/r02o_trace_tests.log
python - <<'PY'
import sys, torch
sys.path.insert(0, ".")
import sdf_viewer_b200 as S
BB = ((-1.0, -1.0, -1.0), (1.0, 1.0, 1.0))
W, H = 1920, 1080
with S.SDFViewer.new_voxels((512, 512, 512), BB, 2) as v:
    v.set_tape(S.tape.demo_tape()); v.fill_all(); v.commit()
    stream = torch.cuda.ExternalStream(v.stream)
    for name, cam in (("default", S.default_camera(W, H)), ("closeup", S.look_at_camera((0.9, 1.1, 1.8), (0, 0, 0), W, H)),
                      ("along -z", S.look_at_camera((0.1, 0.05, 4.0), (0, 0, 0), W, H))):
        for dv in (0, 1):
            v.set_option("trace_distance_volume", dv)
            out = []
            for ss in (65536, 64, 32, 24, 16, 8, 4, 0):
                v.set_option("trace_spec_start", ss)
                for _ in range(3): v.trace_device(cam, W, H)
                v.sync()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                for _ in range(30): v.trace_device(cam, W, H)
                e1.record(stream); v.sync(); torch.cuda.synchronize()
                out.append(f"{ss}: {e0.elapsed_time(e1) / 30:.4f}")
            print(f"{name} dist_volume={dv} ms by spec_start -> " + "  ".join(out), flush=True)
PY
