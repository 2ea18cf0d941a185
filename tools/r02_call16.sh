#!/bin/bash
# round 2 experiment, N = 1: the tile trace kernel with its registers capped (more CTAs per SM)
for mb in 0 20 24 32; do
SDFGPU_TRACE_MB=$mb python - $mb <<'PY'
import sys, torch
sys.path.insert(0, ".")
import sdf_viewer_b200 as S
BB = ((-1.0, -1.0, -1.0), (1.0, 1.0, 1.0))
W, H = 1920, 1080
with S.SDFViewer.new_voxels((512, 512, 512), BB, 2) as v:
    v.set_tape(S.tape.demo_tape()); v.fill_all(); v.commit()
    stream = torch.cuda.ExternalStream(v.stream)
    out = []
    for name, cam in (("default", S.default_camera(W, H)), ("closeup", S.look_at_camera((0.9, 1.1, 1.8), (0, 0, 0), W, H)),
                      ("along -z", S.look_at_camera((0.1, 0.05, 4.0), (0, 0, 0), W, H))):
        for _ in range(3): v.trace_device(cam, W, H)
        v.sync()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(30): v.trace_device(cam, W, H)
        e1.record(stream); v.sync(); torch.cuda.synchronize()
        out.append(f"{name} {e0.elapsed_time(e1) / 30:.4f}")
    print(f"minBlocks {sys.argv[1]}: " + "  ".join(out), flush=True)
PY
done
