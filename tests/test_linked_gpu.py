"""Linked slabs (include/sdfgpu.h "multi-GPU: linked slabs") on ONE device: a group shards the grid along z over
several handles that all live on GPU 0, so the driver's single-GPU box runs the whole protocol -- one-launch
fill (halo slices filled by their holder, or boundary-first with the flag-ordered halo push), the collective
update / resample_box / reset, and the exact ray-hand-off trace in rounds -- and checks it against ONE handle holding the whole grid: volumes bit for bit, frames bit for
bit (RGBA8, depth, G-buffer but for the normals).  tests/multi_gpu_check.py runs the same over real peers.

The reference has no multi-GPU path (src/app/scene/sdf/mod.rs:174); what is pinned here is that sharding changes
nothing the reference's scene (scene/mod.rs:158-225) would see."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

BB = ((-1.0, -1.0, -1.0), (1.0, 1.0, 1.0))


def cameras(S, w, h):
    return [S.default_camera(w, h), S.look_at_camera((0.3, 0.2, 1.4), (0.0, 0.1, 0.0), w, h),
            S.look_at_camera((0.2, 0.1, 0.3), (1, 0.2, -0.4), w, h),   # camera inside the box
            S.look_at_camera((0.1, 0.05, -2.6), (0, 0, 0), w, h),      # looks along +z: every ray crosses every slab
            S.look_at_camera((0.1, 0.05, 2.6), (0, 0, 0), w, h)]       # along -z


def same_frame(got, want, slab_faces, dims):
    g8, gd, gg = got
    w8, wd, wg = want
    assert np.array_equal(g8, w8), f"RGBA8 differs in {(g8 != w8).any(axis=-1).sum()} pixels"
    assert np.array_equal(gd.view(np.uint32), np.clip(wd, 0, 1).view(np.uint32)), "depth differs"
    if gg is None:
        return
    # G-buffer: position, code, raw samples, step count bit-equal; normals (12..14) only away from slab faces (their
    # taps reach lod * D / |dims| < 4 voxels from the hit, beyond the one-slice halo; dead code for the RGBA frame)
    cols = [0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 15]
    a, b = gg[..., cols].view(np.uint32), wg[..., cols].view(np.uint32)
    assert np.array_equal(a, b), f"G-buffer differs in {(a != b).any(axis=-1).sum()} pixels"
    hit = wg[..., 3] >= 0
    z = (wg[..., 2] - BB[0][2]) / (BB[1][2] - BB[0][2]) * dims[2]
    far = hit & np.all([np.abs(z - f) > 6.0 for f in slab_faces] or [np.ones_like(hit)], axis=0)
    n_g, n_w = gg[..., 12:15][far], wg[..., 12:15][far]
    assert np.array_equal(n_g.view(np.uint32), n_w.view(np.uint32)), "normals away from slab faces differ"


@pytest.mark.parametrize("n,dims,wait_mode,halo_push", [(2, (48, 40, 36), 0, False), (2, (48, 40, 36), 0, True), (3, (40, 36, 50), 0, True),
                                                        (2, (32, 32, 33), 1, True), (2, (32, 32, 33), 1, False), (4, (36, 32, 33), 0, False),
                                                        (4, (36, 32, 33), 0, True), (5, (33, 31, 64), 0, False)])
def test_group_equals_single_handle(S, n, dims, wait_mode, halo_push):
    """halo_push False (default): every rank fills its halo slices itself; True: the neighbours push them."""
    w, h = 200, 150
    sdf = S.SDFDemo()
    with S.SDFViewer.new_voxels(dims, BB, 3) as whole:
        if wait_mode:
            # the spin-wait fallback (drivers without stream memory operations) must be chosen before linking, and
            # group_create links at once: build the link by hand -- slabs, option, export, attach
            ranks = []
            for r in range(n):
                zr = (r * dims[2] // n, (r + 1) * dims[2] // n)
                v = S.SDFViewer.new_voxels(dims, BB, 3, z_range=zr)
                v.set_option("link_wait_mode", 1)
                ranks.append(v)
            blobs = [v.link_export(r, n, w, h, gbuf=True, halo_push=halo_push) for r, v in enumerate(ranks)]
            for v in ranks:
                v.link_attach(blobs)
            assert all(v.get_info("linked") == 1 and v.get_info("link_memops") == 0 for v in ranks)
            assert all(v.get_info("link_halo_push") == int(halo_push) and v.get_info("link_trace_stream") == 0 for v in ranks)
            try:
                run_by_hand(S, ranks, whole, sdf, dims, w, h)
            finally:
                for v in ranks:
                    v.sync()
                for v in ranks:
                    v.link_detach()
                for v in ranks:
                    v.close()
            return
        with S.SDFViewerGroup.new_voxels(dims, BB, 3, [0] * n, w, h, gbuf=True, halo_push=halo_push) as g:
            assert g.size == n and g.dims == dims
            assert all(r.get_info("link_halo_push") == int(halo_push) and r.get_info("link_trace_stream") == 0 for r in g.ranks)
            group_body(S, whole, g, sdf, dims, w, h)


def group_body(S, whole, g, sdf, dims, w, h):
    """What a group must reproduce of ONE handle `whole` (both freshly created with 3 loading passes); also run over
    real peers by tests/group_devices_check.py."""
    faces = [r.z_begin for r in g.ranks[1:]]
    # pass by pass: volumes and frames (lod 4, 2, 1; NEAREST then LINEAR) equal the single handle's
    sdf2 = S.SDFDemo()
    for k in range(3):
        it_w = whole.update(sdf, max_passes=1)
        it_g = g.update(sdf2, max_passes=1)
        assert it_w == it_g
        whole.commit(); g.commit()
        t0, t1 = whole.download()
        g0, g1 = g.download()
        assert np.array_equal(t0.view(np.uint32), g0.view(np.uint32)), f"tex0 differs after pass {k}"
        assert np.array_equal(t1.view(np.uint32), g1.view(np.uint32)), f"tex1 differs after pass {k}"
        for cam in cameras(S, w, h):
            want8, want_d = whole.trace_rgba8(cam, w, h)
            _, _, want_g = whole.trace(cam, w, h, gbuf=True)
            same_frame(g.trace(cam, w, h), (want8, want_d, want_g), faces, dims)
    assert g.loading_state()[0] == 0
    # a frame without G-buffer through the rgba8 entry point, several frames back to back (key-frame parity)
    for cam in cameras(S, w, h) * 2:
        want8, want_d = whole.trace_rgba8(cam, w, h)
        same_frame(g.trace_rgba8(cam, w, h) + (None,), (want8, want_d, None), faces, dims)
    # halo slices hold the neighbours' boundary slices
    for r in g.ranks:
        check_halos(r, t0, t1, dims)
    # fill_all (one launch per rank, boundary tiles first) after a reset
    g.reset(1); whole.reset(1)
    g.set_tape(sdf.tape()); whole.set_tape(sdf.tape())
    g.fill_all(); whole.fill_all()
    g.commit(); whole.commit()
    t0, t1 = whole.download()
    g0, g1 = g.download()
    assert np.array_equal(t0.view(np.uint32), g0.view(np.uint32)) and np.array_equal(t1.view(np.uint32), g1.view(np.uint32))
    for r in g.ranks:
        check_halos(r, t0, t1, dims)
    cam = cameras(S, w, h)[3]
    same_frame(g.trace_rgba8(cam, w, h) + (None,), whole.trace_rgba8(cam, w, h) + (None,), faces, dims)
    # dirty boxes: inside one slab, across a face, touching nothing
    for box in ((-0.3, -0.2, -0.9, 0.4, 0.3, -0.7), (-0.5, -0.5, -0.5, 0.5, 0.5, 0.5), (3, 3, 3, 4, 4, 4),
                (-1, -1, -1, 1, 1, 1)):
        other = S.tape.demo_tape() if box[0] == 3 else S.tape.csg_tape(S.tape.csg_primitive_table(12))
        g.set_tape(other); whole.set_tape(other)
        assert g.resample_box(box, count=True) == whole.resample_box(box, count=True)
        t0, t1 = whole.download()
        g0, g1 = g.download()
        assert np.array_equal(t0.view(np.uint32), g0.view(np.uint32)) and np.array_equal(t1.view(np.uint32), g1.view(np.uint32))
        for r in g.ranks:
            check_halos(r, t0, t1, dims)
        # few primitives -> long steps: rays leap over whole slabs and out of the box (hand-offs that skip a rank,
        # positions whose mirrored taps belong to a rank the ray has already left)
        for c in cameras(S, w, h):
            want8, want_d = whole.trace_rgba8(c, w, h)
            _, _, want_g = whole.trace(c, w, h, gbuf=True)
            same_frame(g.trace(c, w, h), (want8, want_d, want_g), faces, dims)


def check_halos(v, t0, t1, dims):
    """The stored slices of a rank (own + halo) equal the whole grid's, bit for bit."""
    import torch
    from sdf_viewer_b200.sharded import _DevMem
    v.sync()
    n = dims[0] * dims[1] * 4
    p0, p1 = v.device_ptrs()
    for p, want in ((p0, t0), (p1, t1)):
        t = torch.as_tensor(_DevMem(p, (v.z_hi - v.z_lo) * n, "<f4"), device=torch.device("cuda", v.get_info("device")))
        got = t.cpu().numpy().reshape(v.z_hi - v.z_lo, dims[1], dims[0], 4)
        assert np.array_equal(got.view(np.uint32), np.ascontiguousarray(want[v.z_lo:v.z_hi]).view(np.uint32)), \
            f"stored slices [{v.z_lo},{v.z_hi}) of the rank owning [{v.z_begin},{v.z_end}) differ (halo exchange)"


def run_by_hand(S, ranks, whole, sdf, dims, w, h):
    """The collective calls issued rank by rank on hand-linked handles (what one process per GPU does, here in one
    process): every rank's call is enqueued before the next rank's, which the flag protocol must tolerate."""
    faces = [r.z_begin for r in ranks[1:]]
    for v in ranks:
        v.set_tape(sdf.tape())
    whole.set_tape(sdf.tape())
    whole.update(None); whole.commit()
    for v in ranks:
        v.update(None)
        v.commit()
    t0, t1 = whole.download()
    for v in ranks:
        v.sync()
    for v in ranks:
        a0, a1 = v.download()
        assert np.array_equal(a0.view(np.uint32), t0[v.z_begin:v.z_end].view(np.uint32))
        assert np.array_equal(a1.view(np.uint32), t1[v.z_begin:v.z_end].view(np.uint32))
        check_halos(v, t0, t1, dims)
    # rank-major issue of a frame needs every rank's rounds in flight at once: only possible from several threads
    # (or processes); here each rank traces in its own thread, the presenter collects
    import threading
    for cam in cameras(S, w, h):
        want8, want_d = whole.trace_rgba8(cam, w, h)
        _, _, want_g = whole.trace(cam, w, h, gbuf=True)
        out, errs = [None], []

        def work(r, v):
            try:
                res = v.trace_linked(cam, w, h, gbuf=True, presenter=(r == 0))
                if r == 0:
                    out[0] = res
            except Exception as e:  # noqa: BLE001
                errs.append(e)

        ts = [threading.Thread(target=work, args=(r, v)) for r, v in enumerate(ranks)]
        for t in ts:
            t.start()
        for t in ts:
            t.join(60)
        assert not errs, errs
        same_frame(out[0], (want8, want_d, want_g), faces, dims)


def test_linked_handle_rejects_solo_calls(S):
    """Entry points that cannot be collective fail loudly on a linked handle; frames larger than the link's fail."""
    dims = (32, 32, 32)
    with S.SDFViewerGroup.new_voxels(dims, BB, 1, [0, 0], 64, 48) as g:
        g.set_tape(S.tape.demo_tape())
        g.fill_all(); g.commit()
        cam = S.default_camera(64, 48)
        with pytest.raises(S.SdfGpuError) as e:
            g.ranks[1].trace_device(cam, 64, 48)
        assert e.value.code == -4
        with pytest.raises(S.SdfGpuError):
            g.trace_rgba8(S.default_camera(128, 96), 128, 96)
        with pytest.raises(S.SdfGpuError):  # no G-buffer frame was reserved
            g.trace(cam, 64, 48, gbuf=True)
        r8, d = g.trace_rgba8(cam, 64, 48)  # and the group still works
        assert (d < 1).any()
    with pytest.raises(S.SdfGpuError):
        S.SDFViewerGroup.new_voxels((8, 8, 3), BB, 1, [0] * 4, 64, 48)  # fewer slices than devices


def test_group_of_one_and_from_bb(S):
    """device_mask with one bit: the plain handle behind the group surface (SDFViewer::from_bb, scene/sdf/mod.rs:46-68)."""
    with S.SDFViewerGroup.from_bb(BB, 40, 2, device_mask=1, max_width=160, max_height=120) as g, \
            S.SDFViewer.from_bb(BB, 40, 2) as v:
        assert g.size == 1 and g.dims == v.dims == (40, 40, 40)
        sdf = S.SDFDemo()
        assert g.update(sdf) == v.update(S.SDFDemo())
        g.commit(); v.commit()
        cam = S.default_camera(160, 120)
        a, b = g.trace_rgba8(cam, 160, 120), v.trace_rgba8(cam, 160, 120)
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
        g0, _ = g.download()
        v0, _ = v.download()
        assert np.array_equal(g0.view(np.uint32), v0.view(np.uint32))
