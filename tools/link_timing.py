"""Development tool (torchrun, N ranks): where the linked trace spends its time.  SDFGPU_LINK_TIMING=1 makes the
library print, per synchronising frame and rank, the microseconds of every round's wait and kernel."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
import sdf_viewer_b200 as S
from sdf_viewer_b200.sharded import ShardedViewer

BB = ((-1.0, -1.0, -1.0), (1.0, 1.0, 1.0))


def main():
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    side = int(sys.argv[1]) if len(sys.argv) > 1 else 512
    halo_push = bool(int(sys.argv[2])) if len(sys.argv) > 2 else False
    trace_mode = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    dims = [side, side, side]
    k, axis = world, 2
    while k > 1:
        dims[axis] *= 2; k //= 2; axis = (axis - 1) % 3
    W, H = 1920, 1080
    sv = ShardedViewer(dims, BB, 2, rank=rank, world=world, device=local, group=dist if world > 1 else None, max_width=W, max_height=H,
                       halo_push=halo_push, trace_mode=trace_mode)
    if rank == 0:
        print(f"== link: halo_push {halo_push}, trace {'stream' if world > 1 and sv.viewer.get_info('link_trace_stream') else 'rounds'}", flush=True)
    v = sv.viewer
    v.set_tape(S.tape.demo_tape())
    sv.fill_all(); sv.commit()
    stream = torch.cuda.ExternalStream(v.stream, device=torch.device("cuda", local))
    for name, cam in (("default", S.default_camera(W, H)), ("closeup", S.look_at_camera((0.9, 1.1, 1.8), (0, 0, 0), W, H)),
                      ("along -z", S.look_at_camera((0.1, 0.05, 4.0), (0, 0, 0), W, H)), ("along x", S.look_at_camera((4.0, 0.3, 0.2), (0, 0, 0), W, H))):
        for _ in range(3):
            sv.trace_host(cam, W, H)  # prints the timing lines
        v.sync(); torch.cuda.synchronize()
        if world > 1: dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(20):
            sv.trace_device(cam, W, H)
        e1.record(stream)
        v.sync(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        if world > 1:
            t = torch.tensor([ms], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX); ms = t.item()
        if rank == 0:
            print(f"== {name}: {ms:.3f} ms per frame, back to back (world {world}, dims {dims})", flush=True)
        if world == 1:
            for variant in (0, 2):
                v.set_option("trace_variant", variant)
                v.trace_device(cam, W, H); v.sync()
                e0.record(stream)
                for _ in range(20):
                    v.trace_device(cam, W, H)
                e1.record(stream); v.sync(); torch.cuda.synchronize()
                print(f"   variant {variant}: {e0.elapsed_time(e1) / 20:.3f} ms", flush=True)
            v.set_option("trace_variant", 0)
    # the bench step (fill + commit + trace, enqueued back to back), per rank
    cam = S.default_camera(W, H)
    for _ in range(5):
        sv.fill_all(); sv.commit(); sv.trace_device(cam, W, H)
    v.sync(); torch.cuda.synchronize()
    if world > 1: dist.barrier()
    n = 30
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(n)]
    for e in evs:
        e[0].record(stream); sv.fill_all(); sv.commit(); e[1].record(stream); sv.trace_device(cam, W, H); e[2].record(stream)
    v.sync(); torch.cuda.synchronize()
    fill = sum(e[0].elapsed_time(e[1]) for e in evs[5:]) / (n - 5)
    trace = sum(e[1].elapsed_time(e[2]) for e in evs[5:]) / (n - 5)
    step = evs[5][0].elapsed_time(evs[-1][2]) / (n - 5)
    print(f"== step loop rank {rank}: fill {fill:.3f} ms, trace {trace:.3f} ms, step {step:.3f} ms", flush=True)
    sv.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
