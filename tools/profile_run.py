"""Small fixed workload for ncu: N fills + N traces of the bench configuration.  Development tool."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import sdf_viewer_b200 as S

side = int(sys.argv[1]) if len(sys.argv) > 1 else 512
workload = sys.argv[2] if len(sys.argv) > 2 else "demo"
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
BB = ((-1.0, -1.0, -1.0), (1.0, 1.0, 1.0))
with S.SDFViewer.from_bb(BB, side, 2) as v:
    v.set_tape(S.tape.demo_tape() if workload == "demo" else S.tape.csg_tape())
    cam = S.default_camera(1920, 1080)
    if os.environ.get("SDFGPU_TRACE_DIST"):  # distance source of the march: 1 dense array, 2 TMU point, 3 TMU linear
        v.set_option("trace_distance_volume", int(os.environ["SDFGPU_TRACE_DIST"]))
    for _ in range(reps):
        v.fill_all()
        v.commit()
        v.trace_device(cam, 1920, 1080)
    v.sync()
print("done")
