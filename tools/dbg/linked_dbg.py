import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import sdf_viewer_b200 as S
BB = ((-1.0, -1.0, -1.0), (1.0, 1.0, 1.0))
n, dims = 3, (40, 36, 50)
w, h = 200, 150
sdf = S.SDFDemo()
cams = [S.default_camera(w, h), S.look_at_camera((0.1, 0.05, -2.6), (0, 0, 0), w, h), S.look_at_camera((0.1, 0.05, 2.6), (0, 0, 0), w, h)]
with S.SDFViewer.new_voxels(dims, BB, 1) as whole, S.SDFViewerGroup.new_voxels(dims, BB, 1, [0] * n, w, h, gbuf=True) as g:
    g.set_tape(sdf.tape()); whole.set_tape(sdf.tape())
    g.fill_all(); whole.fill_all(); g.commit(); whole.commit()
    print("slabs", [(r.z_begin, r.z_end) for r in g.ranks])
    for box in ((-0.3, -0.2, -0.9, 0.4, 0.3, -0.7), (-0.5, -0.5, -0.5, 0.5, 0.5, 0.5), (3, 3, 3, 4, 4, 4), (-1, -1, -1, 1, 1, 1)):
        other = S.tape.demo_tape() if box[0] == 3 else S.tape.csg_tape(S.tape.csg_primitive_table(12))
        g.set_tape(other); whole.set_tape(other)
        print("box", box, g.resample_box(box, count=True), whole.resample_box(box, count=True))
        t0, t1 = whole.download(); g0, g1 = g.download()
        print(" volumes equal", np.array_equal(t0.view(np.uint32), g0.view(np.uint32)), np.array_equal(t1.view(np.uint32), g1.view(np.uint32)))
        for ci, cam in enumerate(cams):
            for rep in range(2):
                g8, gd, gg = g.trace(cam, w, h)
                w8, wd = whole.trace_rgba8(cam, w, h)
                _, _, wg = whole.trace(cam, w, h, gbuf=True)
                bad = (g8 != w8).any(axis=-1)
                print(f"  cam {ci} rep {rep}: rgba8 differs in {bad.sum()} px; depth differs {(gd.view(np.uint32) != np.clip(wd,0,1).view(np.uint32)).sum()}; gbuf code differs {(gg[...,3] != wg[...,3]).sum()} steps differ {(gg[...,15] != wg[...,15]).sum()}")
                if bad.sum():
                    ys, xs = np.nonzero(bad)
                    print("   rows", ys.min(), ys.max(), "cols", xs.min(), xs.max())
                    for y, x in list(zip(ys, xs))[:6]:
                        print("   px", y, x, "got", g8[y, x], gg[y, x, [0, 1, 2, 3, 15]], "want", w8[y, x], wg[y, x, [0, 1, 2, 3, 15]],
                              "z idx", (wg[y, x, 2] + 1) / 2 * dims[2])
