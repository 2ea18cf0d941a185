import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import sdf_viewer_b200 as S
from tools.configs_run import timed, BB
W, H = 1920, 1080
for side in (64, 128, 256, 512):
    with S.SDFViewer.from_bb(BB, side, 2) as v:
        stream = torch.cuda.ExternalStream(v.stream)
        v.set_tape(S.tape.demo_tape()); v.fill_all(); v.commit()
        for name, cam in (("default", S.default_camera(W, H)), ("closeup", S.look_at_camera((0.9, 1.1, 1.8), (0, 0, 0), W, H))):
            for ms_ in (256, 64, 32, 16):
                v.set_option("trace_max_steps", ms_)
                r, d, g = v.trace(cam, W, H, gbuf=True)
                steps = g[..., 15].sum()
                ms = timed(v, stream, lambda: v.trace_device(cam, W, H), 20)
                print(f"side {side} {name} max_steps {ms_}: {ms:.3f} ms  steps {steps:.3e}  {steps/ms/1e6:.1f} Gsteps/s", flush=True)
