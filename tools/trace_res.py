"""Development tool: trace time against the number of rays (frame size) on one volume -- is the kernel bound by
throughput (time ~ rays) or by the latency chain of its longest rays (time ~ constant)?"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import sdf_viewer_b200 as S
BB = ((-1.0, -1.0, -1.0), (1.0, 1.0, 1.0))
side = int(sys.argv[1]) if len(sys.argv) > 1 else 512
with S.SDFViewer.from_bb(BB, side, 1) as v:
    v.set_tape(S.tape.demo_tape()); v.fill_all(); v.commit()
    stream = torch.cuda.ExternalStream(v.stream)
    for (W, H) in ((240, 135), (480, 270), (960, 540), (1920, 1080), (3840, 2160)):
        for name, cam in (("default", S.default_camera(W, H)), ("closeup", S.look_at_camera((0.9, 1.1, 1.8), (0, 0, 0), W, H))):
            row = []
            for variant, dv in ((0, 0), (2, 0), (0, 1)):
                v.set_option("trace_variant", variant); v.set_option("trace_distance_volume", dv)
                v.trace_device(cam, W, H); v.sync()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                for _ in range(30):
                    v.trace_device(cam, W, H)
                e1.record(stream); v.sync(); torch.cuda.synchronize()
                row.append(f"v{variant}/d{dv} {e0.elapsed_time(e1) / 30:.4f}")
            print(f"{side}^3 {W}x{H} {name}: " + "  ".join(row) + " ms", flush=True)
