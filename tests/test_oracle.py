"""Pins the CPU oracle (oracle/sdf_oracle.cpp): hand-derived known answers from the reference's
source lines, the committed golden fixtures, and internal consistency
(tape interpreter == direct restatement; OpenMP fill == reference-order fill).

The reference holds no value-level vectors for this path, and cannot be run here: "parity unpinned"."""
import os

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
BB = ((-1.0, -1.0, -1.0), (1.0, 1.0, 1.0))
f32 = np.float32


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def test_constants(oracle):
    L = oracle.lib()
    assert L.orc_air_dist() == f32(0.1) + f32(0.001234)      # scene/sdf/mod.rs:42
    assert abs(float(L.orc_air_dist()) - 0.101234004) < 1e-9
    lut = (oracle.C.c_float * 256)()
    L.orc_srgb_lut(lut)
    lut = np.array(lut, np.float32)
    assert lut[0] == 0 and lut[255] == 1 and np.all(np.diff(lut) > 0)
    c = f32(10) / f32(255)
    assert lut[10] == c / f32(12.92)                           # linear segment below 0.04045
    assert abs(lut[128] - ((128 / 255 + 0.055) / 1.055) ** 2.4) < 1e-7
    # `(c * 255.0) as u8`: truncation, saturation, NaN -> 0
    for v, want in ((0.0, 0), (1.0, 255), (0.5, 127), (0.6, 153), (0.7, 178), (2.0, 255), (-1.0, 0), (float("nan"), 0),
                    (150 / 255, 150), (24 / 255, 24), (10 / 255, 10), (56 / 255, 56), (70 / 255, 70), (60 / 255, 60)):
        assert L.orc_f32_to_u8(f32(v)) == want, v


def test_demo_known_answers(oracle):
    """SURVEY 8c, derived by hand from demo/mod.rs:58-73, cube.rs:79-89,181-222, sphere.rs:37-47."""
    d_box_corner = f32(1.0) - f32(0.95)
    s = oracle.demo_sample([[1, 1, 1], [-1, -1, -1]])
    for row in s:  # corner: box distance 0.05, sphere is air (d_sph = sqrt(3)-1.05 > 0.1), brick lands on cement
        assert row[0] == d_box_corner
        assert tuple(row[1:]) == (f32(56) / f32(255), f32(70) / f32(255), f32(60) / f32(255), f32(0.4), f32(0.5), f32(1.0))
    # seam at p = (1,0,0): |d_box| ~ |d_sph| -> forced material (demo/mod.rs:62-70)
    row = oracle.demo_sample([[1, 0, 0]])[0]
    assert row[0] == max(f32(1) - f32(0.95), -(f32(1) - f32(1.05)))
    assert tuple(row[1:]) == (f32(0.5), f32(0.6), f32(0.7), f32(0.5), f32(0.0), f32(0.0))
    # centre region: deep inside the removed sphere -> distance = -d_sph
    p = f32(0.015873075)
    row = oracle.demo_sample([[p, p, p]])[0]
    d_sph = np.sqrt(p * p + p * p + p * p, dtype=np.float32) - f32(1.05)
    assert row[0] == -d_sph and abs(row[0] - 1.0225) < 1e-3
    # distance_only skips every material
    s = oracle.demo_sample(np.random.default_rng(1).uniform(-1, 1, (500, 3)), distance_only=True)
    full = oracle.demo_sample(np.random.default_rng(1).uniform(-1, 1, (500, 3)))
    assert np.array_equal(bits(s[:, 0]), bits(full[:, 0]))


def test_voxel_positions_and_store_rules(oracle):
    """scene/sdf/mod.rs:179-182 (three roundings) and :196-208 (clamp, grey, sRGB LUT, occlusion)."""
    v = oracle.Viewer(BB, (64, 64, 64), 2)
    for i, want in ((1, -0.96825397), (31, -0.015873015), (0, -1.0), (63, 1.0)):
        x = v.voxel_pos(i, 0, 0)[0]
        assert x == f32(f32(f32(i) / f32(63)) * f32(2)) + f32(-1)
        assert abs(x - want) < 1e-7
    its = v.update(oracle.Sampler())
    assert its == 294912 and v.len() == 0                       # 32^3 + 64^3 (loading.rs:80-89)
    t0, t1 = v.tex0, v.tex1
    lut = (oracle.C.c_float * 256)(); oracle.lib().orc_srgb_lut(lut)
    assert t0[0, 0, 0, 0] == f32(0.1) + (f32(1.0) - f32(0.95))
    assert tuple(t0[0, 0, 0, 1:]) == (lut[56], lut[70], lut[60])
    assert tuple(t1[0, 0, 0]) == (f32(0.4), f32(0.5), f32(1.0), f32(oracle.lib().orc_air_dist()))
    assert t0[32, 32, 32, 0] == 1.0                             # clamp(0.1 + 1.02, 0, 1)
    assert np.all(t1[..., 3] == f32(oracle.lib().orc_air_dist()))  # tex1.a is never written (:205-208)
    assert t0[..., 0].min() >= 0 and t0[..., 0].max() <= 1
    assert np.all(t1[..., 2] > 0)                               # occlusion <= 0 stored as 1 (:208)


def test_tape_interpreter_equals_direct_demo(oracle, S):
    rng = np.random.default_rng(5)
    pts = np.concatenate([rng.uniform(-1.2, 1.2, (20000, 3)), np.array([[0, 0, 0], [1, 1, 1], [0.95, 0, 0]])]).astype(f32)
    for kw in ({}, dict(cube_half_side=0.7, sphere_radius=0.8), dict(disable_sphere=True),
               dict(cube_material=S.tape.MAT_NORMAL, sphere_material=S.tape.MAT_BRICK), dict(max_distance_custom_material=0.3)):
        P = oracle.demo_params(**{k: (int(v) if isinstance(v, bool) else v) for k, v in kw.items()})
        with np.errstate(all="ignore"):
            a, b = oracle.demo_sample(pts, P), oracle.tape_sample(S.tape.demo_tape(**kw), pts)
        assert np.array_equal(bits(a), bits(b)), kw


def test_parallel_fill_equals_reference_order(oracle, S):
    tape = S.tape.csg_tape(S.tape.csg_primitive_table(25, seed=2))
    a = oracle.Viewer(BB, (20, 17, 9), 3); a.update(oracle.Sampler(tape=tape))
    b = oracle.Viewer(BB, (20, 17, 9), 3); b.fill_all(oracle.Sampler(tape=tape), threads=4)
    assert np.array_equal(bits(a.tex0), bits(b.tex0)) and np.array_equal(bits(a.tex1), bits(b.tex1))


def test_golden_demo_samples(oracle):
    g = np.load(os.path.join(GOLD, "demo_samples.npz"))
    with np.errstate(all="ignore"):
        assert np.array_equal(bits(oracle.demo_sample(g["points"])), bits(g["samples"]))
        assert np.array_equal(bits(oracle.demo_sample(g["points"], distance_only=True)), bits(g["samples_distance_only"]))


def test_golden_volume_and_csg(oracle, S):
    g = np.load(os.path.join(GOLD, "demo_volume_16.npz"))
    v = oracle.Viewer(BB, (16, 16, 16), 2)
    assert v.update(oracle.Sampler(tape=S.tape.demo_tape())) == int(g["iterations"])
    assert np.array_equal(bits(v.tex0), bits(g["tex0"])) and np.array_equal(bits(v.tex1), bits(g["tex1"]))
    g = np.load(os.path.join(GOLD, "csg_samples.npz"))
    table = S.tape.csg_primitive_table(40, seed=11)
    assert np.array_equal(table.tobytes(), g["table"].tobytes()), "csg_primitive_table is no longer deterministic"
    assert np.array_equal(bits(oracle.tape_sample(S.tape.csg_tape(table), g["points"])), bits(g["samples"]))


def test_golden_trace(oracle, S):
    g = np.load(os.path.join(GOLD, "trace_32_160x120.npz"))
    w, h = 160, 120
    rays = S.camera_rays(S.default_camera(w, h), w, h)
    for k in ("origin", "base", "dx", "dy", "bvp"):
        assert np.array_equal(np.array(getattr(rays, k), f32), g[k].astype(f32)), k
    v = oracle.Viewer(BB, (32, 32, 32), 2)
    v.update(oracle.Sampler(tape=S.tape.demo_tape()))
    P = oracle.trace_params(rays, BB, (32, 32, 32), lod=1.0, filter_linear=1)
    rgba, depth, gbuf = oracle.trace(P, v.tex0, v.tex1, w, h, threads=4)
    assert np.array_equal(bits(gbuf), bits(g["gbuf"]))
    assert np.array_equal(bits(depth), bits(g["depth"]))
    np.testing.assert_allclose(rgba, g["rgba"], rtol=1e-6, atol=1e-7)
    hit = gbuf[..., 3] >= 0
    assert 0.05 < hit.mean() < 0.5
    # lighting with one ambient light: rgb = occlusion * albedo * (1 - metallic), then ACES + sRGB
    s0, s1 = gbuf[hit][:, 4:8], gbuf[hit][:, 8:12]
    lin = s1[:, 2:3] * s0[:, 1:4] * (1 - s1[:, 0:1])
    aces = np.clip(lin * (2.51 * lin + 0.03) / (lin * (2.43 * lin + 0.59) + 0.14), 0, 1)
    srgb = np.where(aces < 0.0031308, aces * 12.92, 1.055 * aces ** (1 / 2.4) - 0.055)
    np.testing.assert_allclose(rgba[hit][:, :3], srgb, rtol=1e-4, atol=1e-6)
    assert np.all(rgba[hit][:, 3] == 1) and np.all(rgba[~hit] == 0) and np.all(depth[~hit] == 1)


def test_trace_edge_semantics(oracle, S):
    """Rays that miss the box get code -3; rays whose entry point is within 0.2 of leaving start at
    camera + 0.2 dir and go out of bounds at once (material.frag:136-139,106-109)."""
    g = np.load(os.path.join(GOLD, "trace_32_160x120.npz"))
    codes = g["gbuf"][..., 3]
    assert (codes == -3).mean() > 0.5
    assert (codes == -2).sum() > 0
    assert np.all(g["gbuf"][..., 15][codes == -3] == 0)
    assert g["gbuf"][..., 15].max() <= 255


def test_trace_known_answers_derived_by_hand(oracle):
    """tests/trace_kats.py: hit / miss, w, step count and end position of the centre ray computed on paper from
    material.frag:27-36,97-126 and the GL filter rules -- independent of how the oracle was written."""
    import trace_kats as K
    rays = oracle.camera_rays(K.EYE, K.TARGET, K.UP, K.FOVY, K.W, K.H)
    for name, dims, r, lod, linear, _exp in K.cases():
        t0, t1 = K._volume(dims, r)
        P = oracle.trace_params(rays, K.BB, dims, lod=lod, filter_linear=linear)
        rgba, depth, gbuf = oracle.trace(P, t0, t1, K.W, K.H)
        K.check(name, gbuf, depth, rgba)
    for name, eye, dims, r, exp in K.cases_oracle_only():  # the step limit; the ray origin with the camera inside the box
        t0, t1 = K._volume(dims, r)
        target = (eye[0], eye[1], eye[2] - 1.0)
        P = oracle.trace_params(oracle.camera_rays(eye, target, K.UP, K.FOVY, K.W, K.H), K.BB, dims, lod=1.0, filter_linear=1)
        rgba, depth, gbuf = oracle.trace(P, t0, t1, K.W, K.H)
        K.check_expectation(name, exp, gbuf, depth, rgba)


def test_oracle_equals_an_independent_numpy_restatement(oracle):
    """tests/demo_numpy_ref.py -- a second reading of the reference's Rust, written without looking at the C++ oracle --
    against the oracle: SDFDemo::sample bit for bit on random points and on the voxel positions of a grid (where seams,
    brick joints and the air early-outs are hit exactly), for the defaults and for other parameters; the stored
    texels of SDFViewer::update bit for bit except the sRGB decode (glibc's f32 powf: within 5e-7 of the correctly rounded value)."""
    import demo_numpy_ref as R
    f32 = np.float32
    rng = np.random.default_rng(11)
    dims = (40, 36, 33)
    grid = R.voxel_positions(dims, BB).reshape(-1, 3)
    pts = np.concatenate([rng.uniform(-1.2, 1.2, (100000, 3)).astype(f32), grid,
                          np.array([[1, 0, 0], [0.95, 0.95, 0.95], [0, 0, 0], [-1, 1, -1], [0.25, -0.125, 0.0], [-0.0, 0.0, 1.0]], f32)])
    for kw in ({}, {"cube_half_side": 0.8, "sphere_radius": 0.9, "max_distance_custom_material": 0.1}):
        want = R.demo_sample(pts, kw.get("cube_half_side", 0.95), kw.get("sphere_radius", 1.05), kw.get("max_distance_custom_material", 0.05))
        got = oracle.demo_sample(pts, oracle.demo_params(**kw))
        same = (got.view(np.uint32) == want.view(np.uint32)) | (np.isnan(got) & np.isnan(want))
        assert same.all(), (kw, np.argwhere(~same)[:5], pts[np.argwhere(~same)[:3, 0]], got[~same][:5], want[~same][:5])
        # every branch is exercised: cement, brick, seam, sphere material, air
        assert (want[:, 4] == f32(0.4)).any() and (want[:, 4] == f32(0.2)).any() and (want[:, 4] == f32(0.5)).any()
        assert ((want[:, 1:4] == 0).all(1)).any() and ((want[:, 4] == 0) & (want[:, 1] > 0)).any()
    # positions and stored texels of a whole grid
    o = oracle.Viewer(BB, dims, 1)
    o.fill_all(oracle.Sampler())
    assert np.array_equal(np.array([o.voxel_pos(x, 7, 5) for x in range(dims[0])], f32).view(np.uint32),
                          R.voxel_positions(dims, BB)[5, 7].view(np.uint32))
    t0, t1 = R.store(R.demo_sample(grid), f32(0.1) + f32(0.001234))
    o0, o1 = o.tex0.reshape(-1, 4), o.tex1.reshape(-1, 4)
    assert np.array_equal(t0[:, 0].view(np.uint32), o0[:, 0].view(np.uint32)) and np.array_equal(t1.view(np.uint32), o1.view(np.uint32))
    np.testing.assert_allclose(t0[:, 1:], o0[:, 1:], rtol=5e-7, atol=0)  # glibc's f32 powf against the correctly rounded value


@pytest.mark.parametrize("eye,target,lod,linear", [((2.5, 3.0, 5.0), (0, 0, 0), 1.0, 1), ((0.9, 1.1, 1.8), (0, 0, 0), 1.0, 1),
                                                     ((0.2, 0.1, 0.3), (1, 0.2, -0.4), 1.0, 1), ((2.5, 3.0, 5.0), (0, 0, 0), 2.0, 0),
                                                     ((2.5, 3.0, 5.0), (0, 0, 0), 1.0, 0)])
def test_oracle_trace_equals_an_independent_numpy_restatement(oracle, eye, target, lod, linear):
    """tests/frag_numpy_ref.py -- material.frag and the GL sampling rules read again, in numpy, without looking at the
    oracle or the kernel -- against the oracle's G-buffer: hit / miss class, miss code and step count of (nearly) every
    pixel identical, hit distance and position within 1e-4 (the two differ in operation order, so a grazing ray may
    take a step more or fewer).  Cameras outside and inside the box; LINEAR, NEAREST and the lod-2 snapped fetch."""
    import frag_numpy_ref as R
    dims, w, h = (32, 32, 32), 96, 72
    o = oracle.Viewer(BB, dims, 1)
    o.fill_all(oracle.Sampler())  # (the lod-2 case snaps its fetches on a complete volume: it is the fetch that is compared)
    rays = oracle.camera_rays(eye, target, (0.0, 1.0, 0.0), 45.0, w, h)
    P = oracle.trace_params(rays, BB, dims, lod=lod, filter_linear=linear)
    _, _, g = oracle.trace(P, o.tex0, o.tex1, w, h)
    code, steps, where = R.trace(rays, BB, o.tex0, w, h, lod=lod, linear=bool(linear))
    ocode, osteps = g[..., 3], g[..., 15].astype(np.int32)
    cls = lambda c: np.where(c >= 0, 0, c).astype(np.int32)  # noqa: E731   hit, or which miss
    same_class = cls(code) == cls(ocode)
    assert same_class.mean() > 0.999, (same_class.mean(), np.argwhere(~same_class)[:5])
    assert (ocode >= 0).sum() > 200 and (ocode < 0).sum() > 200
    same_steps = (steps == osteps) & same_class
    assert same_steps.mean() > 0.995, same_steps.mean()
    m = same_steps & (ocode > -3)
    np.testing.assert_allclose(where[m], g[..., 0:3][m], rtol=0, atol=1e-4)
    np.testing.assert_allclose(code[m], ocode[m], rtol=0, atol=1e-4)
    # at the oracle's own hit points: the raw samples of both textures it recorded (:120, :154) and gl_FragDepth (:180-181)
    hit = ocode >= 0
    hp = g[..., 0:3][hit]
    np.testing.assert_allclose(R.sample(o.tex0, hp, BB[0], BB[1], lod, bool(linear)), g[..., 4:8][hit], rtol=1e-5, atol=2e-6)
    np.testing.assert_allclose(R.sample(o.tex1, hp, BB[0], BB[1], lod, bool(linear)), g[..., 8:12][hit], rtol=1e-5, atol=2e-6)
    _, d, _ = oracle.trace(P, o.tex0, o.tex1, w, h)
    np.testing.assert_allclose(R.frag_depth(rays.bvp, hp), d[hit], rtol=1e-5, atol=1e-6)
    assert np.all(d[~hit] == 1.0)
