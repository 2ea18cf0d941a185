"""Trace kernel timing for the four distance sources of the march (option trace_distance_volume):
0 tex0.r in place, 1 dense R32F array, 2 R32F 3-D CUDA array through the TMU in point mode (exact),
3 the same with hardware LINEAR filtering (approximate).  512^3 demo volume, 1920x1080, two cameras.
Also prints mode 3's error against the exact frame.  Run on the GPU box: python tools/trace_modes.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import sdf_viewer_b200 as S
from tools.configs_run import timed, BB

W, H = 1920, 1080
side = int(sys.argv[1]) if len(sys.argv) > 1 else 512
with S.SDFViewer.from_bb(BB, side, 2) as v:
    stream = torch.cuda.ExternalStream(v.stream)
    v.set_tape(S.tape.demo_tape()); v.fill_all(); v.commit()
    for name, cam in (("default", S.default_camera(W, H)), ("closeup", S.look_at_camera((0.9, 1.1, 1.8), (0, 0, 0), W, H))):
        exact = None
        for mode in (0, 1, 2, 3):
            v.set_option("trace_distance_volume", mode)
            r, d, g = v.trace(cam, W, H, gbuf=True)     # builds the distance volume of this mode
            ms = timed(v, stream, lambda: v.trace_device(cam, W, H), 30)
            line = f"side {side} {name} mode {mode}: {ms:.4f} ms  {W * H / ms / 1e6:.2f} Grays/s  steps {g[..., 15].sum():.3e}"
            if mode == 0:
                exact = (r, d, g)
            else:
                same = all(np.array_equal(a.view(np.uint32), b.view(np.uint32)) for a, b in zip(exact, (r, d, g)))
                hit_e, hit = exact[2][..., 3] >= 0, g[..., 3] >= 0
                both = hit_e & hit
                line += f"  bit-identical {same}  hit-mask diff {(hit_e != hit).mean():.2e}"
                if not same:
                    line += (f"  |dt| mean {np.abs(exact[2][..., 3] - g[..., 3])[both].mean():.3e} max "
                             f"{np.abs(exact[2][..., 3] - g[..., 3])[both].max():.3e}  |drgb| max {np.abs(exact[0] - r)[both].max():.3e}")
            print(line, flush=True)
        v.set_option("trace_distance_volume", 0)
