"""A SECOND restatement of the reference's demo SDF and of the store rules of SDFViewer::update, in numpy, written
from the Rust sources alone (not from oracle/sdf_oracle.cpp): /root/reference/src/sdf/demo/mod.rs:51-75,
cube.rs:79-89,164-222, sphere.rs:37-47,122-124, src/sdf/mod.rs:104-126, src/app/scene/sdf/mod.rs:177-208.
tests/test_oracle.py compares the C++ oracle with it bit for bit: two independent readings of the same lines have to
agree before either is trusted.  Every operation is a single f32 operation, in the reference's order (Rust does not
fuse or reassociate)."""
import numpy as np

f32 = np.float32


def _fmod(a, b):
    return np.fmod(a, f32(b)).astype(f32)  # Rust's % on floats: truncated remainder (fmodf); exact


def _tex2d(u, v):
    """compute_tex2d, cube.rs:189-204 -> cement mask"""
    bw, bh = f32(0.5), f32(0.25)
    row_num = v / bh
    brick_offset = np.floor(row_num) / f32(4.0)
    bx = _fmod(np.abs(u + brick_offset), bw)
    by = _fmod(np.abs(v), bh)
    mcd = f32(0.2) / f32(2.0) * bh
    return (bx < mcd) | (bx > bw - mcd) | (by < mcd) | (by > bh - mcd)


def cube_sample(p, half):
    """SDFDemoCube::sample (distance_only = false), brick material -> (n, 7)"""
    x, y, z = p[:, 0], p[:, 1], p[:, 2]
    half = f32(half)
    d = np.maximum(np.maximum(np.abs(x), np.abs(y)), np.abs(z)) - half          # :81
    out = np.zeros((len(p), 7), f32)
    out[:, 0] = d
    tex = ~(d > f32(0.1))                                                          # :83: the air has no texture
    nx = np.where(np.abs(x) > half, np.sign(x), f32(0)).astype(f32)                # normal(), :164-177
    ny = np.where(np.abs(y) > half, np.sign(y), f32(0)).astype(f32)
    nz = np.where(np.abs(z) > half, np.sign(z), f32(0)).astype(f32)
    ax, ay, az = np.abs(nx), np.abs(ny), np.abs(nz)
    # tri-planar choice, :206-220
    c_zy, c_xy, c_zx = _tex2d(z, y), _tex2d(x, y), _tex2d(z, x)
    cement = np.where(ax > ay, np.where(ax > az, c_zy, c_xy), np.where(ay > az, c_zx, c_xy))
    cement_c = np.array([f32(56.) / f32(255.), f32(70.) / f32(255.), f32(60.) / f32(255.)], f32)
    brick_c = np.array([f32(150.) / f32(255.), f32(24.) / f32(255.), f32(10.) / f32(255.)], f32)
    mat = np.where(cement[:, None], np.concatenate([cement_c, [f32(0.4), f32(0.5), f32(1.0)]]).astype(f32),
                   np.concatenate([brick_c, [f32(0.2), f32(0.8), f32(0.0)]]).astype(f32))
    out[tex, 1:] = mat[tex]
    return out


def sphere_sample(p, radius):
    """SDFDemoSphere::sample (distance_only = false), normal material -> (n, 7)"""
    x, y, z = p[:, 0], p[:, 1], p[:, 2]
    mag = np.sqrt((x * x + y * y) + z * z)                                          # cgmath distance / magnitude
    d = mag - f32(radius)                                                          # :39
    out = np.zeros((len(p), 7), f32)
    out[:, 0] = d
    tex = ~(d > f32(0.1))                                                          # :41
    with np.errstate(divide="ignore", invalid="ignore"):
        inv = f32(1.0) / mag                                                       # normalize(): v * (1 / |v|)
        n = np.abs(np.stack([x * inv, y * inv, z * inv], 1)).astype(f32)           # Material::Normal, cube.rs:56
    out[tex, 1:4] = n[tex]
    return out


def demo_sample(points, half=0.95, radius=1.05, seam=0.05):
    """SDFDemo::sample, demo/mod.rs:51-75 -> (n, 7): distance, r, g, b, metallic, roughness, occlusion"""
    p = np.ascontiguousarray(points, f32).reshape(-1, 3)
    box, sph = cube_sample(p, half), sphere_sample(p, radius)
    dist = np.maximum(box[:, 0], -sph[:, 0])                                       # :58
    inter = np.abs(box[:, 0]) - np.abs(sph[:, 0])                                  # :60
    out = np.where((inter < f32(0))[:, None], box, sph)                            # :61
    on_seam = np.abs(inter) <= f32(seam)                                           # :62
    out[on_seam, 1:] = np.array([0.5, 0.6, 0.7, 0.5, 0.0, 0.0], f32)               # :66-69
    out[:, 0] = dist                                                               # :72
    return out.astype(f32)


def voxel_positions(dims, bb):
    """scene/sdf/mod.rs:179-182: index / (size - 1) * bb_size + bb_min, three roundings; (D, H, W, 3)"""
    lo, hi = np.asarray(bb[0], f32), np.asarray(bb[1], f32)
    axes = []
    for a in range(3):
        i = np.arange(dims[a], dtype=f32)
        axes.append(i / f32(dims[a] - 1) * (hi[a] - lo[a]) + lo[a])
    z, y, x = np.meshgrid(axes[2], axes[1], axes[0], indexing="ij")
    return np.stack([x, y, z], -1).astype(f32)


def store(samples, air_dist):
    """scene/sdf/mod.rs:196-208 -> tex0 (n, 4), tex1 (n, 4); the sRGB decode in float64 (the reference's f32 powf is
    compared with a tolerance of 5e-7 by the caller)"""
    s = np.asarray(samples, f32)
    t0 = np.empty((len(s), 4), f32)
    t1 = np.full((len(s), 4), f32(air_dist), f32)
    t0[:, 0] = np.clip(f32(1e-1) + s[:, 0], f32(0), f32(1))                        # :196
    col = s[:, 1:4].copy()
    col[(col == 0).all(1)] = f32(0.5)                                              # :197-200
    u8 = np.clip(np.trunc(col * f32(255.0)), 0, 255).astype(np.uint8)              # Srgba::from: (c * 255) as u8
    c = u8.astype(np.float64) / 255.0
    t0[:, 1:4] = np.where(c < 0.04045, c / 12.92, ((c + 0.055) / 1.055) ** 2.4).astype(f32)  # to_linear_srgb
    t1[:, 0], t1[:, 1] = s[:, 4], s[:, 5]
    t1[:, 2] = np.where(s[:, 6] <= f32(0), f32(1), s[:, 6])                        # :208
    return t0, t1
