"""Host-side logic of the Z-slab sharding (sdf-viewer_b200/sharded.py) on CPU: slab partition,
halo plan, and the halo exchange / key compositing collectives run with world_size 2 and 3 over
`gloo` (the GPU path runs the same functions over NCCL)."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BB = ((-1.0, -1.0, -1.0), (1.0, 1.0, 1.0))


def test_slab_partition_properties(S):
    from sdf_viewer_b200.sharded import slab_range, stored_range, halo_plan
    for depth in (1, 2, 5, 8, 64, 1000, 1024):
        for world in (1, 2, 3, 4, 8):
            ranges = [slab_range(depth, r, world) for r in range(world)]
            assert ranges[0][0] == 0 and ranges[-1][1] == depth
            assert all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))           # contiguous, no overlap
            sizes = [b - a for a, b in ranges]
            assert max(sizes) - min(sizes) <= 1                                       # balanced
            sends, recvs = set(), set()
            for r in range(world):
                zb, ze = ranges[r]
                lo, hi = stored_range(depth, zb, ze)
                for kind, peer, z in halo_plan(depth, r, world):
                    if kind == "send":
                        assert zb <= z < ze
                        sends.add((r, peer, z))
                    else:
                        assert lo <= z < hi and not (zb <= z < ze)                 # lands in a halo slice
                        recvs.add((peer, r, z))
            assert sends == recvs                                                     # every send has its recv


def _worker(rank, world, port, dims, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist
    import orc
    import sdf_viewer_b200 as S
    from sdf_viewer_b200.sharded import slab_range, stored_range, exchange_halos, gather_slabs
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        W, H, D = dims
        tape = S.tape.demo_tape()
        full = orc.Viewer(BB, dims, 1)
        full.fill_all(orc.Sampler(tape=tape), threads=1)
        zb, ze = slab_range(D, rank, world)
        lo, hi = stored_range(D, zb, ze)
        n = W * H * 4
        tex = []
        for src in (full.tex0, full.tex1):
            t = torch.full(((hi - lo) * n,), float("nan"), dtype=torch.float32)
            t[(zb - lo) * n:(ze - lo) * n] = torch.from_numpy(np.ascontiguousarray(src[zb:ze]).reshape(-1))
            tex.append(t)
        n_ops = exchange_halos(dist, tex, dims, rank, world)
        ok = True
        for t, src in zip(tex, (full.tex0, full.tex1)):
            want = torch.from_numpy(np.ascontiguousarray(src[lo:hi]).reshape(-1))
            ok = ok and bool(torch.equal(t.view(torch.int32), want.view(torch.int32)))
        # sort-last compositing: element-wise MIN of (depth bits << 32 | rgba8) keys
        rng = np.random.default_rng(100 + rank)
        depth = rng.uniform(0, 1, 64).astype(np.float32)
        rgba = rng.integers(0, 2 ** 32, 64, dtype=np.uint64)
        keys = (depth.view(np.uint32).astype(np.uint64) << np.uint64(32)) | rgba
        kt = torch.from_numpy(keys.view(np.int64).copy())
        dist.all_reduce(kt, op=dist.ReduceOp.MIN)
        gathered = [torch.zeros(64, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(gathered, torch.from_numpy(keys.view(np.int64).copy()))
        want = np.minimum.reduce([g.numpy().view(np.uint64) for g in gathered])
        ok = ok and np.array_equal(kt.numpy().view(np.uint64), want)
        # exact trace: the distance channel of the whole grid replicated by one broadcast per non-empty slab
        dist_full = np.ascontiguousarray(full.tex0[..., 0]).reshape(-1)
        mine = torch.full((W * H * D,), float("nan"), dtype=torch.float32)
        mine[zb * W * H:ze * W * H] = torch.from_numpy(dist_full[zb * W * H:ze * W * H])
        n_bcast = gather_slabs(dist, mine, dims, world)
        ok = ok and n_bcast == sum(1 for r in range(world) if slab_range(D, r, world)[0] < slab_range(D, r, world)[1])
        ok = ok and bool(torch.equal(mine.view(torch.int32), torch.from_numpy(dist_full).view(torch.int32)))
        q.put((rank, ok, n_ops))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,dims", [(2, (8, 6, 10)), (3, (5, 4, 7)), (3, (4, 4, 2))])
def test_halo_exchange_gloo(world, dims):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000) + world
    procs = [ctx.Process(target=_worker, args=(r, world, port, dims, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
    results = sorted(q.get(timeout=5) for _ in range(world))
    assert [r[0] for r in results] == list(range(world))
    assert all(r[1] for r in results), results
    assert all(p.exitcode == 0 for p in procs)
