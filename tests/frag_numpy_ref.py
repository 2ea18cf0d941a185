"""A SECOND restatement of the reference's fragment shader in numpy, written from
/root/reference/src/app/scene/sdf/material.frag alone (not from oracle/sdf_oracle.cpp or csrc/trace.cu, which one
hand wrote): main() :130-139, sdfRaycast :92-128, sdfSampleRawInterp / Nearest :27-53, sdfOutOfBoundsDist :83-88, and the
OpenGL sampling rules the shader relies on (texel centres at (i + 0.5) / N; GL_LINEAR = the two nearest centres per axis
weighted by the fractional part; GL_MIRRORED_REPEAT; GL_NEAREST = floor(u * N)).  tests/test_oracle.py compares the
oracle's G-buffer with it.  What the rasteriser provides -- `pos`, the point where the pixel's ray enters the bounding
cube (or leaves it, with the camera inside) -- is computed here from the same per-pixel ray basis the oracle gets.

float32 throughout; the operation ORDER is this file's own, so agreement is to a tolerance, not bit for bit."""
import numpy as np

f32 = np.float32


def _mirror(i, n):
    """GL_MIRRORED_REPEAT for integer texel coordinates"""
    m = np.mod(i, 2 * n)
    return np.where(m >= n, 2 * n - 1 - m, m)


def _oob(p, bmin, bmax):
    return np.max(np.maximum(bmin - p, p - bmax), axis=-1)                       # :83-88


def sample(tex, p, bmin, bmax, lod, linear):
    """all four lanes of sdfSampleRawInterp(k, p) at positions p (n, 3) -> (n, 4)"""
    return np.stack([_sample_r(tex[..., c:c + 1], p, np.asarray(bmin, f32), np.asarray(bmax, f32), lod, linear) for c in range(4)], 1)


def frag_depth(bvp, p):
    """gl_FragDepth = (BVP * vec4(hitPos, 1)).z / .w, :180-181; bvp column-major (16,)"""
    m = np.asarray(bvp, f32).reshape(4, 4).T
    q = np.concatenate([p, np.ones((len(p), 1), f32)], 1) @ m.T
    return (q[:, 2] / q[:, 3]).astype(f32)


def _sample_r(tex0, p, bmin, bmax, lod, linear):
    """tex0.r at positions p (n, 3) -- sdfSampleRawInterp(0, p).r"""
    d, h, w = tex0.shape[:3]
    size = np.array([w, h, d], f32)
    p01 = (p - bmin) / (bmax - bmin)                                             # :44 / :30
    if lod != 1.0:                                                               # :31-35
        steps = size / f32(lod)
        p01 = np.floor(p01 * steps + f32(0.5)) / steps                           # GLSL round(): half away from zero, p01 >= 0
    if linear:
        u = p01 * size - f32(0.5)
        i0 = np.floor(u)
        fr = (u - i0).astype(f32)
        i0 = i0.astype(np.int64)
        out = np.zeros(len(p), f32)
        for dz in (0, 1):
            for dy in (0, 1):
                for dx in (0, 1):
                    xi = _mirror(i0[:, 0] + dx, w); yi = _mirror(i0[:, 1] + dy, h); zi = _mirror(i0[:, 2] + dz, d)
                    wgt = (fr[:, 0] if dx else f32(1) - fr[:, 0]) * (fr[:, 1] if dy else f32(1) - fr[:, 1]) * \
                          (fr[:, 2] if dz else f32(1) - fr[:, 2])
                    out += wgt * tex0[zi, yi, xi, 0]
        return out
    i = np.floor(p01 * size).astype(np.int64)
    return tex0[_mirror(i[:, 2], d), _mirror(i[:, 1], h), _mirror(i[:, 0], w), 0]


def trace(rays, bb, tex0, width, height, lod=1.0, linear=True, max_steps=256):
    """-> code (h, w) [w of the hit position: t >= 0 hit, -1 out of steps, -2 out of bounds, -3 no fragment],
    steps (h, w), position (h, w, 3)"""
    bmin, bmax = np.asarray(bb[0], f32), np.asarray(bb[1], f32)
    cam = np.asarray(rays.origin, f32)
    jj, ii = np.meshgrid(np.arange(height, dtype=f32) + f32(0.5), np.arange(width, dtype=f32) + f32(0.5), indexing="ij")
    dirs = (np.asarray(rays.base, f32) + ii[..., None] * np.asarray(rays.dx, f32) + jj[..., None] * np.asarray(rays.dy, f32)).reshape(-1, 3)
    # the fragment: where the pixel's ray meets the cube's nearest front face (camera outside) or its back face (inside)
    with np.errstate(divide="ignore", invalid="ignore"):
        t1, t2 = (bmin - cam) / dirs, (bmax - cam) / dirs
    tn, tf = np.minimum(t1, t2).max(1), np.maximum(t1, t2).min(1)
    has = tf >= np.maximum(tn, 0)
    te = np.where(tn < 0, tf, tn)
    pos = cam + dirs * te[:, None]
    n = len(dirs)
    code = np.full(n, -3.0, f32); steps = np.zeros(n, np.int32); where = np.zeros((n, 3), f32)
    idx = np.flatnonzero(has)
    pos = pos[idx]
    rd = pos - cam
    rd = rd / np.sqrt((rd * rd).sum(1))[:, None]                                  # :134 normalize
    ro = np.where((_oob(pos + rd * f32(0.2), bmin, bmax) > 0)[:, None], cam + rd * f32(0.2), pos)   # :136-139
    t = np.zeros(len(idx), f32)
    alive = np.ones(len(idx), bool)
    c = np.full(len(idx), -1.0, f32); st = np.zeros(len(idx), np.int32)
    for i in range(max_steps):                                                    # :97
        if not alive.any():
            break
        a = np.flatnonzero(alive)
        if i >= max_steps - 1:                                                    # :99-102
            c[a] = -1.0; st[a] = i; alive[a] = False
            break
        out = _oob(ro[a], bmin, bmax) > f32(1e-4)                                 # :106-109
        c[a[out]] = -2.0; st[a[out]] = i; alive[a[out]] = False
        a = a[~out]
        dist = _sample_r(tex0, ro[a], bmin, bmax, lod, linear) - f32(1e-1)        # :112-113, :59
        hit = dist < f32(1e-5)                                                    # :117
        c[a[hit]] = t[a[hit]]; st[a[hit]] = i; alive[a[hit]] = False
        a, dist = a[~hit], dist[~hit]
        t[a] += dist                                                              # :124
        ro[a] += rd[a] * dist[:, None]                                            # :125
    code[idx], steps[idx], where[idx] = c, st, ro
    return code.reshape(height, width), steps.reshape(height, width), where.reshape(height, width, 3)
