"""Run under torchrun with N >= 2 ranks (one GPU each).  First the default path -- slab handles linked through the
C ABI (halo slices filled by their holder or pushed behind flags, exact ray-hand-off trace streamed or in rounds; no
torch.distributed call after set-up) -- checked against
the CPU oracle (volumes, halos) and against ONE handle holding the whole grid (frames, bit for bit); then the
fallbacks (NCCL / IPC halo exchange, sort-last and replicated-volume traces).  tests/test_sharded_gpu.py launches it."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import orc  # noqa: E402
import sdf_viewer_b200 as S  # noqa: E402
from sdf_viewer_b200.sharded import ShardedViewer  # noqa: E402

BB = ((-1.0, -1.0, -1.0), (1.0, 1.0, 1.0))


def check_slab(sv, v, full, dims, rank):
    """Owned slices and exchanged halo slices equal the oracle's full volume, bit for bit."""
    v.sync(); torch.cuda.synchronize(); dist.barrier()
    for t, want in zip(sv._tex, (full.tex0, full.tex1)):
        got = t.cpu().numpy().reshape(v.z_hi - v.z_lo, dims[1], dims[0], 4)
        assert np.array_equal(got.view(np.uint32), np.ascontiguousarray(want[v.z_lo:v.z_hi]).view(np.uint32)), \
            f"rank {rank}: stored slab [{v.z_lo},{v.z_hi}) differs from the oracle (halo exchange, fused={sv.fused})"


def check_linked(rank, world, local, dims, w, h, sdf, full, want_its, halo_push=False, trace_mode=0):
    """The default path (halo slices filled by their holder, one streaming trace kernel per rank) and its variants
    (halo slices pushed by the neighbours; trace in rounds).  After ShardedViewer's set-up nothing below calls
    torch.distributed but the test's own barriers around host-side comparisons."""
    sv = ShardedViewer(dims, BB, 2, rank=rank, world=world, device=local, group=dist, max_width=w, max_height=h, gbuf=True,
                       halo_push=halo_push, trace_mode=trace_mode)
    assert sv.linked, getattr(sv, "link_error", "linking failed")
    v = sv.viewer
    assert v.get_info("linked") == 1 and v.get_info("link_halo_push") == int(halo_push)
    assert v.get_info("link_trace_stream") == (0 if trace_mode == 1 else 1)
    its = sv.update(S.SDFDemo())
    sv.commit()
    assert its == want_its
    check_slab(sv, v, full, dims, rank)
    cams = [S.default_camera(w, h), S.look_at_camera((0.3, 0.2, 1.4), (0.0, 0.1, 0.0), w, h),
            S.look_at_camera((0.2, 0.1, 0.3), (1, 0.2, -0.4), w, h), S.look_at_camera((0.1, 0.05, -2.6), (0, 0, 0), w, h),
            S.look_at_camera((0.1, 0.05, 2.6), (0, 0, 0), w, h)]
    with S.SDFViewer.new_voxels(dims, BB, 2, device=local) as whole:
        whole.set_tape(sdf.tape()); whole.update(None); whole.commit()
        for c in cams * 2:
            got8, got_d = sv.trace_host(c, w, h)
            if rank == 0:
                want8, want_d = whole.trace_rgba8(c, w, h)
                assert np.array_equal(got8, want8), "linked trace differs from the single-volume frame"
                assert np.array_equal(got_d.view(np.uint32), np.clip(want_d, 0, 1).view(np.uint32))
        got8, got_d, got_g = v.trace_linked(cams[3], w, h, gbuf=True, presenter=(rank == 0))
        if rank == 0:
            _, _, want_g = whole.trace(cams[3], w, h, gbuf=True)
            cols = [0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 15]
            assert np.array_equal(got_g[..., cols].view(np.uint32), want_g[..., cols].view(np.uint32)), "linked G-buffer differs"
        # reset + one-launch fill_all, then a dirty box that crosses the slab faces, then one inside rank 0's slab
        sv.reset(1); whole.reset(1)
        v.set_tape(sdf.tape()); whole.set_tape(sdf.tape())
        sv.fill_all(); whole.fill_all()
        sv.commit(); whole.commit()
        check_slab(sv, v, full, dims, rank)
        other = S.tape.csg_tape(S.tape.csg_primitive_table(12))
        for box in ((-0.5, -0.5, -0.6, 0.5, 0.5, 0.6), (-0.3, -0.2, -0.95, 0.4, 0.3, -0.8)):
            v.set_tape(other); whole.set_tape(other)
            sv.resample_box(box); whole.resample_box(box)
            t0, t1 = whole.download()
            v.sync(); torch.cuda.synchronize(); dist.barrier()
            for t, want in zip(sv._tex, (t0, t1)):
                got = t.cpu().numpy().reshape(v.z_hi - v.z_lo, dims[1], dims[0], 4)
                assert np.array_equal(got.view(np.uint32), np.ascontiguousarray(want[v.z_lo:v.z_hi]).view(np.uint32)), \
                    f"rank {rank}: stored slab differs from the single handle after resample_box{box}"
            got8, got_d = sv.trace_host(cams[0], w, h)
            if rank == 0:
                want8, want_d = whole.trace_rgba8(cams[0], w, h)
                assert np.array_equal(got8, want8) and np.array_equal(got_d.view(np.uint32), np.clip(want_d, 0, 1).view(np.uint32))
        # frames enqueued back to back without host synchronisation (the device-timed loop of bench.py)
        for c in cams * 3:
            sv.fill_all()
            sv.trace_device(c, w, h)
        got8, got_d = sv.trace_host(cams[1], w, h)
        if rank == 0:
            whole.fill_all()
            want8, want_d = whole.trace_rgba8(cams[1], w, h)
            assert np.array_equal(got8, want8) and np.array_equal(got_d.view(np.uint32), np.clip(want_d, 0, 1).view(np.uint32))
        # ... and frames with NOTHING between them: the presenter's unpack of frame t must not be locked out by the
        # streaming kernel of frame t + 1, which fills every SM and waits for the ranks that wait for that unpack
        for c in cams * 4:
            sv.trace_device(c, w, h)
        got8, got_d = sv.trace_host(cams[2], w, h)
        if rank == 0:
            want8, want_d = whole.trace_rgba8(cams[2], w, h)
            assert np.array_equal(got8, want8) and np.array_equal(got_d.view(np.uint32), np.clip(want_d, 0, 1).view(np.uint32))
    dist.barrier()
    if rank == 0:
        print(f"linked path ok: world {world}, halo {'pushed' if halo_push else 'filled locally'}, trace "
              f"{'streamed' if v.get_info('link_trace_stream') else 'in rounds'}, memops {v.get_info('link_memops')}, "
              f"{v.get_info('link_round_epoch')} trace launches")
    sv.close()


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dims = (48, 40, 36)
    w, h = 200, 150
    sdf = S.SDFDemo()
    full = orc.Viewer(BB, dims, 2)
    want_its = full.update(orc.Sampler(tape=sdf.tape()))
    check_linked(rank, world, local, dims, w, h, sdf, full, want_its)                                # the default
    check_linked(rank, world, local, dims, w, h, sdf, full, want_its, halo_push=True, trace_mode=2)  # pushed halos
    check_linked(rank, world, local, dims, w, h, sdf, full, want_its, halo_push=True, trace_mode=1)  # ... and rounds
    check_linked(rank, world, local, dims, w, h, sdf, full, want_its, trace_mode=1)
    fused_modes = []
    for fused in (False, True):                # NCCL send/recv exchange, then the fused in-kernel P2P exchange
        sv = ShardedViewer(dims, BB, 2, rank=rank, world=world, device=local, group=dist, fused=fused, linked=False)
        fused_modes.append(sv.fused)
        v = sv.viewer
        its = sv.update(S.SDFDemo())           # both passes + halo exchange
        sv.commit()
        assert its == want_its
        check_slab(sv, v, full, dims, rank)
        if not fused:
            sv.close()
    assert fused_modes[0] is False
    if rank == 0:
        print("fused halo exchange:", "CUDA IPC peer stores" if fused_modes[1] else "unavailable, NCCL send/recv used")
    n = dims[0] * dims[1] * 4
    # a locally recomputed halo equals the exchanged one (pure function of position)
    with S.SDFViewer.new_voxels(dims, BB, 2, device=local, z_range=(v.z_begin, v.z_end)) as v2:
        v2.set_tape(sdf.tape()); v2.fill_all(); v2.sync()
        p0, _ = v2.device_ptrs()
        from sdf_viewer_b200.sharded import _DevMem
        t2 = torch.as_tensor(_DevMem(p0, (v.z_hi - v.z_lo) * n, "<f4"), device=torch.device("cuda", local))
        assert torch.equal(t2.view(torch.int32), sv._tex[0].view(torch.int32)), "exchanged halo != recomputed halo"
    # composited frame == MIN over the oracle's per-slab traces
    cam = S.default_camera(w, h)
    rgba8, depth = sv.trace_host(cam, w, h)
    cmin, cmax, lod, lin = v.trace_params(cam, w, h, slab_clip=True)
    P = orc.trace_params(S.camera_rays(cam, w, h), BB, dims, lod=lod, filter_linear=lin, z_lo=v.z_lo, z_hi=v.z_hi,
                         clip_min=cmin, clip_max=cmax)
    ro, do, _ = orc.trace(P, full.tex0[v.z_lo:v.z_hi], full.tex1[v.z_lo:v.z_hi], w, h, gbuf=False, threads=2)
    d_all = [torch.zeros(h, w) for _ in range(world)]
    dist.all_gather_object(d_all, torch.from_numpy(np.clip(do, 0, 1)))
    want_depth = np.minimum.reduce([d.numpy() for d in d_all])
    np.testing.assert_allclose(depth, want_depth, rtol=1e-5)
    hit = depth < 1
    assert 0.05 < hit.mean() < 0.5 and rgba8[hit][:, 3].min() == 255 and rgba8[~hit].max() == 0
    # exact trace: replicated distance volume (NCCL broadcasts) + owner shading + MIN composite == the frame
    # of ONE handle that holds the whole grid, bit for bit, for several cameras; camera-only frames re-use
    # the gathered volume
    with S.SDFViewer.new_voxels(dims, BB, 2, device=local) as whole:
        whole.set_tape(sdf.tape()); whole.update(None); whole.commit()
        for k, c in enumerate((cam, S.look_at_camera((0.3, 0.2, 1.4), (0.0, 0.1, 0.0), w, h),
                               S.look_at_camera((0.2, 0.1, 0.3), (1, 0.2, -0.4), w, h))):
            want8, want_d = whole.trace_rgba8(c, w, h)
            got8, got_d = sv.trace_exact_host(c, w, h, gather=(k == 0))
            assert np.array_equal(got8, want8), f"rank {rank}: exact sharded trace differs from the single-volume frame"
            assert np.array_equal(got_d.view(np.uint32), np.clip(want_d, 0, 1).view(np.uint32))
    dist.barrier()
    if rank == 0:
        print(f"multi_gpu_check ok: world {world}, dims {dims}, {its} iterations, {int(hit.sum())} hit pixels")
    sv.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
