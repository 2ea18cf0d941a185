"""torchrun -N: breaks the sharded fill time into kernel / barrier / exchange (development tool)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, torch.distributed as dist
import sdf_viewer_b200 as S
from sdf_viewer_b200.sharded import ShardedViewer, exchange_halos
import bench

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
dims = bench.grid_for(world, 512)


def timed(sv, fn, reps=50):
    v = sv.viewer
    for _ in range(3):
        fn()
    v.sync(); torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(sv._stream)
    for _ in range(reps):
        fn()
    e1.record(sv._stream); v.sync(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / reps], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.item()


for fused in (True, False):
    sv = ShardedViewer(dims, bench.BB, 2, rank=rank, world=world, device=local, group=dist, fused=fused)
    sv.viewer.set_tape(S.tape.demo_tape())
    res = {
        "kernel_only": timed(sv, sv.viewer.fill_all),
        "barrier_only": timed(sv, sv._barrier),
        "fill_all": timed(sv, sv.fill_all),
    }
    if not fused:
        def ex():
            with torch.cuda.stream(sv._stream):
                exchange_halos(dist, sv._tex, sv.dims, rank, world)
        res["exchange_only"] = timed(sv, ex)
    if rank == 0:
        print(f"world {world} dims {dims} fused={sv.fused}: " + ", ".join(f"{k} {v:.4f} ms" for k, v in res.items()), flush=True)
    dist.barrier()
    sv.close()
dist.destroy_process_group()
