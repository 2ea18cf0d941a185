"""Known-answer tests for the sphere tracer, derived BY HAND from the reference's shader text
(/root/reference/src/app/scene/sdf/material.frag) and the GL sampling rules -- not from the oracle and not from the
kernel, which were written by the same hand and could share a misreading.  tests/test_oracle.py runs them against
the CPU oracle, tests/test_trace_gpu.py against the CUDA kernels.

Set-up of every case: bounding box [-1, 1]^3, camera at (0, 0, 5) looking at the origin (up = +y), a 3 x 3 frame;
pixel (1, 1) is the centre of the frame, so its ray runs down the z axis: it enters the box at (0, 0, 1) with
direction (0, 0, -1) (rayOrigin = pos, :133-134; pos + 0.2 dir is inside, so :136-139 does not apply).  The
volumes depend on z only, so x / y filtering is inert.

A "step" is one pass of the loop of sdfRaycast (:97-126) that reaches :124; the G-buffer's step count is the value of
`i` when the loop breaks.

KAT 1  constant volume, tex0.r = 0.6 everywhere -> sampleDist = 0.6 - 0.1 = 0.5 (:59; 0.6f - 0.1f rounds to 0.5f).
       z: 1, 0.5, 0, -0.5, -1 are all within 1e-4 of the box (:106), each is sampled and advanced; the sixth position,
       z = -1.5, is out of bounds: break with w = -2 at i = 5, rayPos = (0, 0, -1.5).

KAT 2  LINEAR filter, lod 1, N = 5 slices at z_k = -1 + 2k/4 = -1, -0.5, 0, 0.5, 1, holding what the fill stores for
       the plane d = z - 0.05: tex0.r[k] = clamp(0.1 + z_k - 0.05, 0, 1) = 0, 0, 0.05, 0.55, 1 (scene/sdf/mod.rs:196).
       GL samples texel coordinate u*N - 0.5 with u = (z + 1) / 2, clamped to [0, N - 1] (CLAMP_TO_EDGE):
         i = 0: z = 1     -> 4.5, clamps to texel 4: r = 1, sampleDist = 0.9, z <- 0.1, t = 0.9
         i = 1: z = 0.1   -> 2.25: 0.75 * 0.05 + 0.25 * 0.55 = 0.175, sampleDist = 0.075, z <- 0.025, t = 0.975
         i = 2: z = 0.025 -> 2.0625: 0.9375 * 0.05 + 0.0625 * 0.55 = 0.08125, sampleDist = -0.01875 < 1e-5:
                HIT (:117-121) with w = t = 0.975 at (0, 0, 0.025), i = 2.
       The texel-centre offset of the reference is visible here: the plane is at 0.05, the shader lands on 0.025.

KAT 3  NEAREST filter while loading, lod 2 (:27-36), N = 8 slices, tex0.r = 0.4 (sampleDist 0.3) except ONE solid
       slice with r = 0.05.  roundSteps = 8 / 2 = 4; the fetch reads texel floor(round(4 * p01) / 4 * 8) (clamped to 7):
         z = 1: p01 = 1 -> round(4) = 4 -> 8 -> texel 7;  z = 0.7: 0.85 -> round(3.4) = 3 -> texel 6;
         z = 0.4: 0.7 -> round(2.8) = 3 -> texel 6;        z = 0.1: 0.55 -> round(2.2) = 2 -> texel 4;
         z = -0.2: 0.4 -> round(1.6) = 2 -> texel 4;       z = -0.5: 0.25 -> round(1.0) = 1 -> texel 2;
         z = -0.8: 0.1 -> round(0.4) = 0 -> texel 0;       z = -1.1: out of bounds.
       a) solid slice 5: never read (an un-snapped NEAREST fetch would read it at z = 0.4): miss, w = -2, i = 7.
       b) solid slice 6: hit at z = 0.7, i = 1, w = t = 0.3.
       c) solid slice 4: hit at z = 0.1, i = 3, w = t = 0.9.

Two more, checked against the CPU oracle only (tests/test_oracle.py; the kernels are compared with the oracle in these
regimes by test_trace_gpu.py: cameras inside the box, step counts of every pixel exact):

KAT 4  the step limit (:97-102, maxSteps = 256 at :142): constant volume tex0.r = f32(0.104): sampleDist =
       f32(0.104) - f32(0.1) = 0.0040000007 (exact: the operands are within a factor of two).  Never < 1e-5, never out of
       bounds (1 - 255 * 0.004 = -0.02): iterations i = 0 .. 254 each advance the ray, iteration i = 255 returns w = -1
       before it samples: a MISS with i = 255 at z = 1 - 255 * 0.0040000007 = -0.0200002.

KAT 5  the ray origin with the camera INSIDE the box (:133-139).  Camera (0, 0, 0.5) looking down -z: the front faces
       of the cube are behind the near plane, the fragment of the centre pixel lies on the back face, pos = (0, 0, -1).
       pos + 0.2 dir = (0, 0, -1.2) is outside the box, so rayOrigin = cameraPosition + 0.2 dir = (0, 0, 0.3) (:138).
       Constant volume 0.6 (sampleDist 0.5): z = 0.3, -0.2, -0.7 are sampled, z = -1.2 is out of bounds: miss, w = -2,
       i = 3, rayPos = (0, 0, -1.2).
"""
import numpy as np

BB = ((-1.0, -1.0, -1.0), (1.0, 1.0, 1.0))
EYE, TARGET, UP, FOVY = (0.0, 0.0, 5.0), (0.0, 0.0, 0.0), (0.0, 1.0, 0.0), 45.0
W = H = 3
PIXEL = (1, 1)  # row, column


def _volume(dims, r_of_slice):
    """tex0 / tex1 (D, H, W, 4): distance channel per slice, white colour, metallic 0, roughness 0, occlusion 1."""
    w, h, d = dims
    t0 = np.ones((d, h, w, 4), np.float32)
    t0[..., 0] = np.asarray(r_of_slice, np.float32)[:, None, None]
    t1 = np.zeros((d, h, w, 4), np.float32)
    t1[..., 2] = 1.0
    t1[..., 3] = np.float32(0.1) + np.float32(0.001234)
    return t0, t1


def records(dims, r_of_slice):
    """The same volume as SDFSample records (n, 7) for sdfgpu_ingest_samples: distance = r - 0.1 is stored back as
    clamp(0.1 + d) (scene/sdf/mod.rs:196); colour white, metallic 0, roughness 0, occlusion 1."""
    w, h, d = dims
    dist = (np.asarray(r_of_slice, np.float64) - 0.1).astype(np.float32)
    rec = np.zeros((d, h, w, 7), np.float32)
    rec[..., 0] = dist[:, None, None]
    rec[..., 1:4] = 1.0
    rec[..., 6] = 1.0
    return rec.reshape(-1, 7)


def cases():
    """(name, dims, tex0.r per slice, lod, filter_linear, loading passes to run first, expected)
    expected = dict(hit, code (w of :101/:107/:118), steps, z)"""
    out = []
    out.append(("constant volume leaves through the far face", (4, 4, 4), [0.6] * 4, 1.0, 1,
                dict(hit=False, code=-2.0, steps=5, z=-1.5)))
    z5 = -1.0 + 2.0 * np.arange(5) / 4.0
    out.append(("plane through the LINEAR filter: texel-centre offset", (5, 5, 5), list(np.clip(0.1 + (z5 - 0.05), 0.0, 1.0)), 1.0, 1,
                dict(hit=True, code=0.975, steps=2, z=0.025)))
    for solid, exp in ((5, dict(hit=False, code=-2.0, steps=7, z=-1.1)), (6, dict(hit=True, code=0.3, steps=1, z=0.7)),
                       (4, dict(hit=True, code=0.9, steps=3, z=0.1))):
        r = [0.4] * 8
        r[solid] = 0.05
        out.append((f"lod-2 snapped NEAREST fetch, solid slice {solid}", (8, 8, 8), r, 2.0, 0, exp))
    return out


def check(name, gbuf, depth, rgba):
    """gbuf (3, 3, 16) of the frame; the centre pixel against the expectation."""
    exp = [c for c in cases() if c[0] == name][0][5]
    g = gbuf[PIXEL]
    assert (g[3] >= 0) == exp["hit"], (name, g[3])
    assert int(g[15]) == exp["steps"], (name, "steps", g[15], exp["steps"])
    np.testing.assert_allclose(g[3], exp["code"], rtol=1e-5, atol=1e-6, err_msg=name)
    np.testing.assert_allclose(g[0:3], [0.0, 0.0, exp["z"]], rtol=1e-5, atol=2e-6, err_msg=name)
    if exp["hit"]:
        # white albedo, metallic 0, occlusion 1, ambient 1: lit colour 1 -> tone / colour mapped, alpha = tint alpha
        assert rgba[PIXEL][3] == 1.0 and depth[PIXEL] < 1.0
    else:
        assert not rgba[PIXEL].any() and depth[PIXEL] == 1.0  # :145-149


def cases_oracle_only():
    """(name, eye, dims, tex0.r per slice, expected) -- KAT 4 and KAT 5 of the module docstring."""
    d = float(np.float32(0.104) - np.float32(0.1))
    return [("the step limit: 255 advances, then w = -1", EYE, (4, 4, 4), [0.104] * 4,
             dict(hit=False, code=-1.0, steps=255, z=1.0 - 255.0 * d)),
            ("camera inside the box: the ray starts 0.2 in front of it", (0.0, 0.0, 0.5), (4, 4, 4), [0.6] * 4,
             dict(hit=False, code=-2.0, steps=3, z=-1.2))]


def check_expectation(name, exp, gbuf, depth, rgba):
    g = gbuf[PIXEL]
    assert (g[3] >= 0) == exp["hit"], (name, g[3])
    assert int(g[15]) == exp["steps"], (name, "steps", g[15], exp["steps"])
    np.testing.assert_allclose(g[3], exp["code"], rtol=1e-5, atol=1e-6, err_msg=name)
    np.testing.assert_allclose(g[0:3], [0.0, 0.0, exp["z"]], rtol=1e-5, atol=2e-6, err_msg=name)
    assert not rgba[PIXEL].any() and depth[PIXEL] == 1.0
