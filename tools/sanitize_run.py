"""A small tour of every kernel for compute-sanitizer (memcheck / racecheck): run under gpurun as
`compute-sanitizer --tool memcheck python tools/sanitize_run.py`."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import sdf_viewer_b200 as S

BB = ((-1.0, -1.0, -1.0), (1.0, 1.0, 1.0))
for prog in (1, 2, 3):
    for vpt in (1, 8):
        with S.SDFViewer.new_voxels((33, 17, 10), BB, 2) as v:
            v.set_option("fill_program", prog); v.set_option("fill_voxels_per_thread", vpt)
            v.set_tape(S.tape.demo_tape())
            v.update(None); v.fill_all(); v.resample_box((-0.3, -0.2, -0.5, 0.4, 0.3, 0.2), count=True)
            v.commit(); v.trace(S.default_camera(96, 64), 96, 64, gbuf=True); v.trace_rgba8(S.default_camera(96, 64), 96, 64)
for prog in (1, 3):
    with S.SDFViewer.new_voxels((40, 24, 16), BB, 1) as v:
        v.set_option("fill_program", prog)
        v.set_tape(S.tape.csg_tape(S.tape.csg_primitive_table(300, seed=5)))
        v.fill_all(); v.download()
with S.SDFViewer.new_voxels((16, 16, 16), BB, 3, z_range=(4, 9)) as v:
    v.set_tape(S.tape.demo_tape()); v.update(None, max_passes=1); v.commit()
    k = v.trace_slab_keys(S.default_camera(64, 48), 64, 48); v.keys_download(k, 64, 48)
    pos = v.voxel_positions(16 * 16 * 4, 100); v.ingest_samples(16 * 16 * 4, np.zeros((100, 7), np.float32))
print("sanitize tour done")
