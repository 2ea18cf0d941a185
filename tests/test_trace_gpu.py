"""GPU parity of the sphere tracer against the CPU oracle's restatement of material.frag
(/root/reference/src/app/scene/sdf/material.frag:92-182).  Both sides trace the SAME volume
(downloaded from the GPU fill) with the SAME ray parameters (sdfgpu_camera_rays).

Tolerances (BASELINE.json north_star: 1e-5 relative): hit/miss class and step count exact;
hit position, t, raw tex0/tex1 at the hit, depth: 1e-5 relative (in practice bit-equal, the
arithmetic is unfused f32 on both sides); final RGBA: 1e-5 relative + 1e-6 absolute, because
device powf and glibc powf differ in the last ulp."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

BB = ((-1.0, -1.0, -1.0), (1.0, 1.0, 1.0))
RTOL = 1e-5


def fill(S, tape, dims, passes=2, bb=BB, max_passes=0, commit=True):
    v = S.SDFViewer.new_voxels(dims, bb, passes)
    v.set_tape(tape)
    v.update(None, max_passes=max_passes)
    if commit:
        v.commit()
    return v


def compare(S, oracle, v, cam, w, h, bb=BB, frac_exact=0.999):
    rg, dg, gg = v.trace(cam, w, h, gbuf=True)
    t0, t1 = v.download()
    rays = S.camera_rays(cam, w, h)
    cmin, cmax, lod, lin = v.trace_params(cam, w, h)
    P = oracle.trace_params(rays, bb, v.dims, lod=lod, filter_linear=lin, tone_mapping=cam.tone_mapping,
                            color_mapping=cam.color_mapping, gamma=cam.gamma, tint=list(cam.tint),
                            ambient=list(cam.ambient))
    ro, do, go = oracle.trace(P, t0, t1, w, h)
    # classification and step counts are exact
    assert np.array_equal(gg[..., 3] < 0, go[..., 3] < 0), "hit mask differs"
    miss = go[..., 3] < 0
    assert np.array_equal(gg[..., 3][miss], go[..., 3][miss]), "miss codes differ"
    assert np.array_equal(gg[..., 15], go[..., 15]), "step counts differ"
    hit = ~miss
    assert hit.sum() > 0.05 * w * h, "camera does not see the SDF"
    np.testing.assert_allclose(gg[hit], go[hit], rtol=RTOL, atol=1e-7)
    np.testing.assert_allclose(dg, do, rtol=RTOL, atol=0)
    np.testing.assert_allclose(rg, ro, rtol=RTOL, atol=1e-6)
    assert np.all(rg[miss] == 0) and np.all(dg[miss] == 1.0)
    # in practice the G-buffer is bit-equal
    same = ((gg.view(np.uint32) == go.view(np.uint32)) | (np.isnan(gg) & np.isnan(go))).all(axis=-1)
    assert same.mean() >= frac_exact, f"only {same.mean():.4f} of G-buffer records bit-equal"
    return hit.mean()


def test_default_camera_640x480(S, oracle):
    """BASELINE config 1: demo SDF, 64^3, 2 passes, 640x480, default scene camera."""
    with fill(S, S.tape.demo_tape(), (64, 64, 64)) as v:
        cmin, cmax, lod, lin = v.trace_params(S.default_camera(640, 480), 640, 480)
        assert lod == 1.0 and lin == 1
        cov = compare(S, oracle, v, S.default_camera(640, 480), 640, 480)
        assert 0.05 < cov < 0.5


@pytest.mark.parametrize("eye,target", [((0.9, 1.1, 1.8), (0, 0, 0)), ((0.2, 0.1, 0.3), (1, 0.2, -0.4)),
                                        ((-3.0, 0.4, 0.2), (0, 0, 0)), ((0.0, 4.0, 0.01), (0, 0, 0))])
def test_other_cameras(S, oracle, eye, target):
    """Close-up, camera INSIDE the box (origin = camera + 0.2 dir, material.frag:136-139), axis views."""
    w, h = 320, 200
    with fill(S, S.tape.demo_tape(), (48, 48, 48)) as v:
        compare(S, oracle, v, S.look_at_camera(eye, target, w, h), w, h)


def test_while_loading_lod(S, oracle):
    """After the coarse pass only (lod = 2^passes_left = 4 or 2): nearest-snapped sampling
    (material.frag:27-36) under the NEAREST filter, then under LINEAR once a lod-1 commit happened."""
    w, h = 320, 240
    cam = S.default_camera(w, h)
    v = S.SDFViewer.new_voxels((64, 64, 64), BB, 3)
    with v:
        v.set_tape(S.tape.demo_tape())
        v.update(None, max_passes=1)
        v.commit()
        assert v.trace_params(cam, w, h)[2:] == (4.0, 0)
        compare(S, oracle, v, cam, w, h)
        v.update(None, max_passes=1)
        v.commit()
        assert v.trace_params(cam, w, h)[2:] == (2.0, 0)
        compare(S, oracle, v, cam, w, h)
        v.update(None)
        v.commit()
        assert v.trace_params(cam, w, h)[2:] == (1.0, 1)
        compare(S, oracle, v, cam, w, h)
        # a later change starts a 3-pass re-sample: lod 4 again, but the filter stays LINEAR
        import ctypes as C
        it = C.c_uint64()
        S.viewer.check(v._lib.sdfgpu_update(v._h, (C.c_float * 6)(-1, -1, -1, 1, 1, 1), 1, C.byref(it)), v._h)
        v.commit()
        assert v.trace_params(cam, w, h)[2:] == (4.0, 1)
        compare(S, oracle, v, cam, w, h)


def test_trace_variants_identical(S):
    """The heavy-first 8x8-tile grid (default) and the plain 2-D grid trace the same frame, bit for bit,
    including pixels outside the screen rectangle of the box, odd frame sizes and a camera inside the box."""
    with fill(S, S.tape.demo_tape(), (64, 64, 64)) as v:
        for (w, h, cam) in ((640, 480, None), (333, 211, None), (320, 200, ((0.2, 0.1, 0.3), (1, 0.2, -0.4))),
                            (200, 300, ((4.0, 0.5, 0.2), (0, 3.0, 0)))):
            c = S.default_camera(w, h) if cam is None else S.look_at_camera(cam[0], cam[1], w, h)
            frames = []
            for variant, dist_volume in ((0, 0), (1, 0), (0, 1), (0, 2)):  # 2: the TMU, point mode
                v.set_option("trace_variant", variant)
                v.set_option("trace_distance_volume", dist_volume)  # distance-only copy of tex0.r for the march
                frames.append(v.trace(c, w, h, gbuf=True))
            for other in frames[1:]:
                for a, b in zip(frames[0], other):
                    assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
        # the copy follows every change of the volume
        v.set_tape(S.tape.demo_tape(sphere_radius=0.8))
        v.fill_all()
        c = S.default_camera(320, 240)
        with_copy = v.trace(c, 320, 240, gbuf=True)
        v.set_option("trace_distance_volume", 0)
        without = v.trace(c, 320, 240, gbuf=True)
        for a, b in zip(with_copy, without):
            assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
        v.set_option("trace_variant", 0)


def test_hardware_linear_filter_fast_mode(S, oracle):
    """trace_distance_volume = 3 marches through the texture unit's own trilinear filter (one fetch per
    step).  Its weights have 8 fractional bits, so it is NOT inside the 1e-5 bar and never the default;
    this pins how far off it is: the same pixels hit (up to a sliver at silhouettes) and hit depths agree
    to the filter's quantisation of a voxel."""
    w, h = 320, 240
    dims = (64, 64, 64)
    with fill(S, S.tape.demo_tape(), dims) as v:
        for cam in (S.default_camera(w, h), S.look_at_camera((0.9, 1.1, 1.8), (0, 0, 0), w, h)):
            exact = v.trace(cam, w, h, gbuf=True)
            v.set_option("trace_distance_volume", 3)
            fast = v.trace(cam, w, h, gbuf=True)
            v.set_option("trace_distance_volume", 0)
            hit_e, hit_f = exact[2][..., 3] >= 0, fast[2][..., 3] >= 0
            assert (hit_e != hit_f).mean() < 1e-2
            both = hit_e & hit_f
            assert both.sum() > 1000
            # hit distance along the ray: a small fraction of a voxel for all but grazing rays
            voxel = 2.0 / dims[0]
            dt = np.abs(exact[2][..., 3] - fast[2][..., 3])[both]
            assert np.quantile(dt, 0.9) < voxel / 4
            assert not np.array_equal(exact[0], fast[0])   # it really is a different (approximate) path
            assert np.abs(exact[0] - fast[0])[both].mean() < 0.05


def test_rgba8_frame(S):
    """sdfgpu_trace_rgba8 == round(clamp(RGBA32F frame) * 255), same depth."""
    w, h = 320, 240
    with fill(S, S.tape.demo_tape(), (48, 48, 48)) as v:
        cam = S.default_camera(w, h)
        rf, df, _ = v.trace(cam, w, h)
        r8, d8 = v.trace_rgba8(cam, w, h)
    assert np.array_equal(d8.view(np.uint32), df.view(np.uint32))
    want = np.rint(np.clip(rf, 0, 1) * np.float32(255)).astype(np.uint8)
    assert np.abs(r8.astype(int) - want.astype(int)).max() <= 1 and (r8 != want).mean() < 1e-3


def test_rgba8_frame_in_bands(S):
    """sdfgpu_trace_rgba8 traces the frame in bands of tile rows and copies each band while the next is traced
    (option trace_bands): the same pixels for any number of bands, odd frame sizes, bands above / below / across the
    screen rectangle of the box, more bands than tile rows."""
    with fill(S, S.tape.demo_tape(), (48, 48, 48)) as v:
        for (w, h, cam) in ((640, 480, None), (333, 211, None), (64, 19, None), (320, 200, ((0.2, 0.1, 0.3), (1, 0.2, -0.4))),
                            (200, 300, ((4.0, 0.5, 0.2), (0, 3.0, 0)))):
            c = S.default_camera(w, h) if cam is None else S.look_at_camera(cam[0], cam[1], w, h)
            v.set_option("trace_bands", 1)
            want8, want_d = v.trace_rgba8(c, w, h)
            assert cam is not None or (want_d < 1).any()
            for bands in (2, 3, 6, 7, 32):
                v.set_option("trace_bands", bands)
                got8, got_d = v.trace_rgba8(c, w, h)
                assert np.array_equal(got8, want8) and np.array_equal(got_d.view(np.uint32), want_d.view(np.uint32)), (w, h, bands)
        v.set_option("trace_bands", 6)


def test_tile_order_from_the_previous_frame(S):
    """The tile grid starts the tiles with the longest marches of the PREVIOUS frame first (option trace_tile_order):
    whatever the previous frame was -- the same camera, another one, another frame size, a frame in bands -- the
    pixels are those of the row-by-row order."""
    with fill(S, S.tape.demo_tape(), (64, 64, 64)) as v:
        shots = [(640, 480, None), (640, 480, None), (640, 480, ((0.9, 1.1, 1.8), (0, 0, 0))), (640, 480, None),
                 (333, 211, None), (640, 480, ((0.2, 0.1, 0.3), (1, 0.2, -0.4))), (640, 480, ((0.2, 0.1, 0.3), (1, 0.2, -0.4)))]
        frames = {}
        for order in (0, 2):
            v.set_option("trace_tile_order", order)
            frames[order] = []
            for (w, h, cam) in shots:
                c = S.default_camera(w, h) if cam is None else S.look_at_camera(cam[0], cam[1], w, h)
                frames[order].append(v.trace(c, w, h, gbuf=True) + v.trace_rgba8(c, w, h))
        for a, b in zip(frames[0], frames[2]):
            for x, y in zip(a, b):
                assert np.array_equal(x.view(np.uint8), y.view(np.uint8))
        v.set_option("trace_tile_order", 1)


def test_uncommitted_nearest(S, oracle):
    """Before any commit lod stays 1 and the GL filter is NEAREST (scene/sdf/mod.rs:110-111)."""
    w, h = 256, 192
    with fill(S, S.tape.demo_tape(), (40, 40, 40), commit=False) as v:
        cam = S.default_camera(w, h)
        assert v.trace_params(cam, w, h)[2:] == (1.0, 0)
        compare(S, oracle, v, cam, w, h)


def test_nonuniform_volume_and_tone_modes(S, oracle):
    bb = ((-0.8, -1.0, -0.6), (1.0, 0.7, 0.9))
    w, h = 200, 160
    v = S.SDFViewer.from_bb(bb, 56, 2)
    with v:
        v.set_tape(S.tape.csg_tape(S.tape.csg_primitive_table(60, seed=3)))
        v.fill_all()
        v.commit()
        for tone, cmap, gamma in ((2, 1, 0.0), (1, 0, 0.0), (3, 1, 2.2), (0, 0, 0.0)):
            cam = S.look_at_camera((2.0, 1.5, 2.5), (0.1, -0.1, 0.1), w, h)
            cam.tone_mapping, cam.color_mapping, cam.gamma = tone, cmap, gamma
            cam.tint[:] = [0.9, 0.8, 1.0, 0.75]
            compare(S, oracle, v, cam, w, h, bb=bb)


def test_slab_keys_composite(S, oracle):
    """Sort-last multi-GPU trace: per-slab key images, MIN-composited, equal the oracle's per-slab
    traces composited the same way; and match the single-volume frame except within 2/255."""
    dims = (48, 48, 48)
    w, h = 240, 180
    tape = S.tape.demo_tape()
    cam = S.default_camera(w, h)
    cuts = [0, 16, 32, 48]
    keys = []
    with fill(S, tape, dims) as full:
        rg, dg, _ = full.trace(cam, w, h)
        t0, t1 = full.download()
    for zb, ze in zip(cuts[:-1], cuts[1:]):
        with S.SDFViewer.new_voxels(dims, BB, 2, z_range=(zb, ze)) as v:
            v.set_tape(tape)
            v.update(None)
            v.commit()
            k = v.trace_slab_keys(cam, w, h)
            rgba8, depth = v.keys_download(k, w, h)
            cmin, cmax, lod, lin = v.trace_params(cam, w, h, slab_clip=True)
            P = oracle.trace_params(S.camera_rays(cam, w, h), BB, dims, lod=lod, filter_linear=lin, z_lo=v.z_lo,
                                    z_hi=v.z_hi, clip_min=cmin, clip_max=cmax)
            ro, do, _ = oracle.trace(P, t0[v.z_lo:v.z_hi], t1[v.z_lo:v.z_hi], w, h, gbuf=False)
            np.testing.assert_allclose(depth, np.clip(do, 0, 1), rtol=RTOL)
            want8 = np.rint(np.clip(ro, 0, 1) * 255).astype(np.uint8)
            assert np.abs(rgba8.astype(int) - want8.astype(int)).max() <= 1
            keys.append((depth.view(np.uint32).astype(np.uint64) << np.uint64(32)) | rgba8.view(np.uint32)[..., 0])
    comp = np.minimum.reduce(keys)
    comp_depth = (comp >> np.uint64(32)).astype(np.uint32).view(np.float32)
    comp_rgba = (comp & np.uint64(0xffffffff)).astype(np.uint32).view(np.uint8).reshape(h, w, 4)
    # sort-last restarts rays at slab faces: same surface, hit point may differ slightly
    full8 = np.rint(np.clip(rg, 0, 1) * 255).astype(np.uint8)
    agree = (np.abs(comp_rgba.astype(int) - full8.astype(int)).max(axis=-1) <= 2)
    assert agree.mean() > 0.98
    hit = dg < 1
    assert np.median(np.abs(comp_depth[hit] - dg[hit])) < 1e-4


@pytest.mark.parametrize("ranges", [((0, 14), (14, 28)), ((0, 9), (9, 10), (10, 10), (10, 28))])
@pytest.mark.parametrize("state", ["linear", "nearest", "loading"])
def test_exact_sharded_trace_equals_single_volume(S, oracle, ranges, state):
    """The exact multi-GPU trace on one device: slab handles (one per would-be rank) each hold a copy of the
    whole grid's distance channel -- gathered here through the host-staging calls instead of NCCL -- march every
    ray through it and shade only the hits they own; the element-wise MIN of their keys is the RGBA8 + depth
    frame of a single handle that holds the whole grid, bit for bit (same step sequence, same texels).
    Covers a one-slice slab, an empty slab, the LINEAR / NEAREST filter states and a half-loaded volume
    (lod 2 snapping)."""
    dims = (32, 24, 28)
    w, h = 200, 150
    tape = S.tape.demo_tape()

    def load(v):
        v.set_tape(tape)
        if state == "loading":
            v.update(None, max_passes=1)    # only the coarse pass: lod = 2, NEAREST filter state
        else:
            v.fill_all()
        if state != "nearest":
            v.commit()

    cams = [S.default_camera(w, h), S.look_at_camera((0.3, 0.2, 1.4), (0.0, 0.1, 0.0), w, h),
            S.look_at_camera((0.2, 0.1, 0.3), (1, 0.2, -0.4), w, h)]
    with S.SDFViewer.new_voxels(dims, BB, 2) as single:
        load(single)
        want = [single.trace_rgba8(c, w, h) for c in cams]
    slabs = [S.SDFViewer.new_voxels(dims, BB, 2, z_range=r) for r in ranges]
    try:
        parts = []
        for v in slabs:
            load(v)
            _, first, count = v.exact_trace_prepare()
            parts.append((first, v.dist_volume_read(first, count)))
        for v in slabs:                                   # the "all-gather"
            for first, values in parts:
                v.dist_volume_write(first, values)
        for cam, (want8, want_depth) in zip(cams, want):
            keys = []
            for v in slabs:
                rgba8, depth = v.keys_download(v.trace_exact_keys(cam, w, h), w, h)
                keys.append((depth.view(np.uint32).astype(np.uint64) << np.uint64(32)) | rgba8.view(np.uint32)[..., 0])
            best = np.minimum.reduce(keys)
            got8 = (best & np.uint64(0xffffffff)).astype(np.uint32).view(np.uint8).reshape(h, w, 4)
            got_depth = (best >> np.uint64(32)).astype(np.uint32).view(np.float32)
            assert np.array_equal(got8, want8)
            assert np.array_equal(got_depth.view(np.uint32), np.clip(want_depth, 0, 1).view(np.uint32))
            # every hit is shaded by exactly one slab
            hits = sum(((k >> np.uint64(32)).astype(np.uint32).view(np.float32) < 1.0).astype(int) for k in keys)
            assert hits.max() <= 1 and (hits == 1).sum() == (want_depth < 1.0).sum() > 0
        # the volume changed: the trace refuses to run on a stale distance volume
        slabs[0].fill_all()
        with pytest.raises(S.SdfGpuError):
            slabs[0].trace_exact_keys(cams[0], w, h)
    finally:
        for v in slabs:
            v.close()


def test_gl_presenter_needs_a_gl_context(S):
    """sdfgpu_gl_register / sdfgpu_trace_gl (CUDA <-> GL interop presenter): on a box without a current GL
    context registration fails loudly with SDFGPU_ERR_CUDA, trace_gl without a registration is a state
    error, and the handle keeps working through the host-buffer path."""
    from sdf_viewer_b200 import _lib
    w, h = 64, 48
    with fill(S, S.tape.demo_tape(), (24, 24, 24)) as v:
        cam = S.default_camera(w, h)
        lib = v._lib
        assert lib.sdfgpu_trace_gl(v._h, cam) == _lib.SDFGPU_ERR_STATE
        assert b"sdfgpu_gl_register" in lib.sdfgpu_last_error(v._h)
        assert lib.sdfgpu_gl_register(v._h, 0, 0, 0x0DE1, w, h) == _lib.SDFGPU_ERR_INVALID
        rc = lib.sdfgpu_gl_register(v._h, 1, 0, 0x0DE1, w, h)      # texture 1 of a GL context that does not exist
        assert rc == _lib.SDFGPU_ERR_CUDA, rc
        assert b"GL context" in lib.sdfgpu_last_error(v._h)
        assert lib.sdfgpu_trace_gl(v._h, cam) == _lib.SDFGPU_ERR_STATE
        assert lib.sdfgpu_gl_unregister(v._h) == 0
        r8, d = v.trace_rgba8(cam, w, h)
        assert (d < 1).any() and r8[d < 1][:, 3].min() == 255


def test_trace_known_answers_derived_by_hand(S):
    """tests/trace_kats.py on the CUDA kernels (every trace variant): expectations computed on paper from
    material.frag:27-36,97-126, not by the oracle.  The volumes go in through sdfgpu_ingest_samples."""
    import trace_kats as K
    cam = S.look_at_camera(K.EYE, K.TARGET, K.W, K.H, up=K.UP, fovy_deg=K.FOVY)
    for name, dims, r, lod, linear, _exp in K.cases():
        passes = 1 if lod == 1.0 else 2
        with S.SDFViewer.new_voxels(dims, K.BB, passes) as v:
            v.set_tape(S.tape.demo_tape())
            v.update(None, max_passes=1)            # lod 1: loaded; lod 2: the coarse pass of two
            v.ingest_samples(0, K.records(dims, r))  # then the whole volume is replaced by the KAT's
            v.commit()
            assert v.trace_params(cam, K.W, K.H)[2:] == (lod, linear), name
            for variant in (0, 1, 2):
                v.set_option("trace_variant", variant)
                rgba, depth, gbuf = v.trace(cam, K.W, K.H, gbuf=True)
                K.check(name, gbuf, depth, rgba)
