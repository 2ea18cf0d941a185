"""Development tool: how much of the trace time is the tail of the longest rays?  Times the frame with maxSteps
(material.frag:142) lowered -- frames differ, this is a timing experiment only -- and prints the step histogram."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import sdf_viewer_b200 as S
BB = ((-1.0, -1.0, -1.0), (1.0, 1.0, 1.0))
W, H = 1920, 1080
with S.SDFViewer.from_bb(BB, 512, 1) as v:
    v.set_tape(S.tape.demo_tape()); v.fill_all(); v.commit()
    stream = torch.cuda.ExternalStream(v.stream)
    for name, cam in (("default", S.default_camera(W, H)), ("closeup", S.look_at_camera((0.9, 1.1, 1.8), (0, 0, 0), W, H))):
        _, _, g = v.trace(cam, W, H, gbuf=True)
        steps = g[..., 15][g[..., 3] != -3]
        hist = np.bincount(np.minimum(steps.astype(int) // 16, 16))
        print(name, "rays entering", len(steps), "steps/16 histogram", hist.tolist(), flush=True)
        for variant in (0, 2):
            v.set_option("trace_variant", variant)
            for ms_ in (256, 128, 96, 64, 48, 32, 16):
                v.set_option("trace_max_steps", ms_)
                v.trace_device(cam, W, H); v.sync()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                for _ in range(30):
                    v.trace_device(cam, W, H)
                e1.record(stream); v.sync(); torch.cuda.synchronize()
                print(f"  variant {variant} maxSteps {ms_:3d}: {e0.elapsed_time(e1) / 30:.4f} ms", flush=True)
            v.set_option("trace_max_steps", 256)
        # a frame of only the rectangle's pixels vs the whole frame: cost of the outside pixels
    v.set_option("trace_variant", 0)
