"""Times the BASELINE.json configurations that are not the bench line (development tool, run under gpurun):
C3 CSG-1k 512^3, C5 dirty-block re-sample at 512^3, progressive update passes, trace step statistics."""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import sdf_viewer_b200 as S

BB = ((-1.0, -1.0, -1.0), (1.0, 1.0, 1.0))


def timed(v, stream, fn, reps):
    fn(); v.sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(reps):
        fn()
    e1.record(stream); v.sync(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    which = sys.argv[1:] or ["csg", "dirty", "passes", "trace"]
    side = 512
    if "csg" in which:
        with S.SDFViewer.from_bb(BB, side, 2) as v:
            stream = torch.cuda.ExternalStream(v.stream)
            v.set_tape(S.tape.csg_tape())
            for prog in (3, 1):
                for vpt in (8, 4, 2):
                    v.set_option("fill_program", prog); v.set_option("fill_voxels_per_thread", vpt)
                    ms = timed(v, stream, v.fill_all, 3)
                    print(f"C3 csg-1k {side}^3 prog={prog} vpt={vpt}: {ms:.3f} ms  {side**3/ms/1e6:.2f} Gsamples/s", flush=True)
            v.set_option("fill_program", 0); v.set_option("fill_voxels_per_thread", 0)
            v.fill_all(); v.commit()
            for (w, h) in ((1920, 1080),):
                cam = S.default_camera(w, h)
                ms = timed(v, stream, lambda: v.trace_device(cam, w, h), 10)
                print(f"C3 trace {w}x{h}: {ms:.3f} ms {w*h/ms/1e6:.2f} Grays/s", flush=True)
    if "dirty" in which or "passes" in which:
        with S.SDFViewer.from_bb(BB, side, 2) as v:
            stream = torch.cuda.ExternalStream(v.stream)
            sdf = S.SDFDemo()
            t = time.perf_counter(); it = v.update(sdf); v.sync(); dt = time.perf_counter() - t
            print(f"initial load 2 passes (conditional kernels, incl. JIT compile): {it} iterations in {dt*1e3:.1f} ms wall", flush=True)
            v.reset(2)
            ms = timed(v, stream, lambda: (v.reset(2), v.update(None)), 5)
            print(f"reset + update(2 passes) {side}^3: {ms:.3f} ms", flush=True)
            # C5: parameter change -> whole bbox dirty (as the reference demo reports), 3-pass re-sample
            def sweep():
                sdf.set_parameter("sphere_radius", 1.05 if sdf.params["sphere_radius"] < 1.05 else 1.04)
                v.update(sdf)
            ms = timed(v, stream, sweep, 5)
            print(f"C5 full-bbox change: set_tape + 3-pass conditional update: {ms:.3f} ms  ({1e3/ms:.0f} Hz)", flush=True)
            # C5 fast path: dirty AABB of 128^3 voxels
            h = 128 / 511.0
            box = (-h, -h, -h, h, h, h)
            n = v.resample_box(box, count=True)
            ms = timed(v, stream, lambda: v.resample_box(box), 20)
            print(f"C5 resample_box {n} voxels (~128^3): {ms*1e3:.1f} us  {n/ms/1e6:.2f} Gsamples/s", flush=True)
            box = (-1, -1, -1, 1, 1, 1)
            ms = timed(v, stream, lambda: v.resample_box(box), 5)
            print(f"C5 resample_box whole grid: {ms:.3f} ms", flush=True)
    if "trace" in which:
        with S.SDFViewer.from_bb(BB, side, 2) as v:
            stream = torch.cuda.ExternalStream(v.stream)
            v.set_tape(S.tape.demo_tape()); v.fill_all(); v.commit()
            for name, cam in (("default", S.default_camera(1920, 1080)), ("closeup", S.look_at_camera((0.9, 1.1, 1.8), (0, 0, 0), 1920, 1080))):
                r, d, g = v.trace(cam, 1920, 1080, gbuf=True)
                steps = g[..., 15]; code = g[..., 3]
                entered = code != -3
                print(f"trace stats {name}: box-entering rays {entered.mean():.3f}, hits {(code>=0).mean():.3f}, "
                      f"steps mean(entered) {steps[entered].mean():.1f} p50 {np.percentile(steps[entered],50):.0f} "
                      f"p90 {np.percentile(steps[entered],90):.0f} p99 {np.percentile(steps[entered],99):.0f} max {steps.max():.0f}, "
                      f"out-of-steps {(code==-1).sum()}", flush=True)
                for variant, dv in ((1, 0), (0, 0), (0, 1)):
                  v.set_option("trace_variant", variant); v.set_option("trace_distance_volume", dv)
                  ms = timed(v, stream, lambda: v.trace_device(cam, 1920, 1080), 20)
                  print(f"trace {name} variant={variant} dist_volume={dv} 1920x1080 LINEAR: {ms:.3f} ms {1920*1080/ms/1e6:.2f} Grays/s; total steps {steps.sum():.3e} -> {steps.sum()/ms/1e6:.1f} Gsteps/s", flush=True)


if __name__ == "__main__":
    main()
