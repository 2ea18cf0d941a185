"""Times the fill kernel variants (voxels/thread x CTAs/SM x store policy) and the trace; prints a table.
Development tool: run under gpurun."""
import sys, os, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import sdf_viewer_b200 as S

BB = ((-1.0, -1.0, -1.0), (1.0, 1.0, 1.0))


def time_fill(v, stream, reps=10):
    v.fill_all(); v.fill_all(); v.sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(reps):
        v.fill_all()
    e1.record(stream)
    v.sync(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    side = int(sys.argv[1]) if len(sys.argv) > 1 else 512
    workload = sys.argv[2] if len(sys.argv) > 2 else "demo"
    tape = S.tape.demo_tape() if workload == "demo" else S.tape.csg_tape()
    rows = []
    with S.SDFViewer.from_bb(BB, side, 2) as v:
        stream = torch.cuda.ExternalStream(v.stream)
        v.set_tape(tape)
        nvox = side ** 3
        for prog, vpt, ctas, streaming in [(p, a, b, c) for p in (1, 2, 3) for a in (1, 2, 4, 8) for b in (0, 1, 2, 3) for c in (1,)]:
            if workload != "demo" and prog == 2:
                continue
            v.set_option("fill_program", prog)
            v.set_option("fill_voxels_per_thread", vpt)
            v.set_option("fill_ctas_per_sm", ctas)
            try:
                ms = time_fill(v, stream, reps=5 if workload == "demo" else 2)
            except Exception as e:
                print("fail", vpt, ctas, e)
                continue
            rows.append((ms, vpt, ctas, streaming, prog))
            print(f"prog={prog} vpt={vpt} ctas={ctas}: {ms:.3f} ms  {nvox/ms/1e6:.1f} Gsamples/s  {nvox*32/ms/1e6:.0f} GB/s", flush=True)
        rows.sort()
        print("best:", rows[:5])
        ms, vpt, ctas, streaming, prog = rows[0]
        v.set_option("fill_program", prog); v.set_option("fill_voxels_per_thread", vpt); v.set_option("fill_ctas_per_sm", ctas)
        v.fill_all(); v.commit()
        for (w, h) in ((640, 480), (1920, 1080), (3840, 2160)):
            for name, cam in (("default", S.default_camera(w, h)), ("closeup", S.look_at_camera((0.9, 1.1, 1.8), (0, 0, 0), w, h))):
                v.trace_device(cam, w, h); v.sync()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                for _ in range(10):
                    v.trace_device(cam, w, h)
                e1.record(stream); v.sync(); torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / 10
                print(f"trace {w}x{h} {name}: {ms:.3f} ms  {w*h/ms/1e6:.2f} Grays/s", flush=True)


if __name__ == "__main__":
    main()
