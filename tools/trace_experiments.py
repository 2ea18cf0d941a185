import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import sdf_viewer_b200 as S
from tools.configs_run import timed, BB
W, H = 1920, 1080
for side in (512,):
    with S.SDFViewer.from_bb(BB, side, 2) as v:
        stream = torch.cuda.ExternalStream(v.stream)
        v.set_tape(S.tape.demo_tape()); v.fill_all(); v.commit()
        for name, cam in (("default", S.default_camera(W, H)), ("closeup", S.look_at_camera((0.9, 1.1, 1.8), (0, 0, 0), W, H))):
            for ms_ in (256, 128, 64, 48, 32):
                v.set_option("trace_max_steps", ms_)
                r, d, g = v.trace(cam, W, H, gbuf=True)
                steps = g[..., 15]
                ms = timed(v, stream, lambda: v.trace_device(cam, W, H), 20)
                print(f"side {side} {name} max_steps {ms_}: {ms:.3f} ms  steps {steps.sum():.3e} rays>=max-1: {(steps >= ms_ - 1).sum()}  {steps.sum()/ms/1e6:.1f} Gsteps/s", flush=True)
        v.set_option("trace_max_steps", 256)
        # empty frame cost: camera looking away from the box
        cam = S.look_at_camera((2.5, 3.0, 5.0), (5.0, 6.0, 10.0), W, H)
        ms = timed(v, stream, lambda: v.trace_device(cam, W, H), 20)
        print(f"looking away (no ray enters the box): {ms:.3f} ms", flush=True)
