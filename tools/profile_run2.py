"""Round-2 workloads for ncu (development tool): `mesh` = fill + marching cubes at 512^3; `linked` = a 2-slab group on one
device, fill + exact trace (the round kernel); `points` = the tape in point mode."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import sdf_viewer_b200 as S

what = sys.argv[1] if len(sys.argv) > 1 else "mesh"
side = int(sys.argv[2]) if len(sys.argv) > 2 else 512
BB = ((-1.0, -1.0, -1.0), (1.0, 1.0, 1.0))
if what == "mesh":
    with S.SDFViewer.from_bb(BB, side, 1) as v:
        v.set_tape(S.tape.demo_tape()); v.fill_all()
        for _ in range(2):
            print(v.mesh(download=False))
elif what == "linked":
    W, H = 1920, 1080
    with S.SDFViewerGroup.new_voxels((side, side, side), BB, 1, [0, 0], W, H) as g:
        g.set_tape(S.tape.demo_tape())
        cam = S.default_camera(W, H)
        for _ in range(3):
            g.fill_all(); g.commit()
            g.trace_rgba8(cam, W, H)
elif what == "points":
    with S.SDFViewer.from_bb(BB, 64, 1) as v:
        v.set_tape(S.tape.demo_tape())
        pts = (np.random.default_rng(1).random((1 << 22, 3), np.float32) * 2 - 1).astype(np.float32)
        for _ in range(2):
            v.sample_points(pts)
print("done")
