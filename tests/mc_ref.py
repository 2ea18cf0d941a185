"""CPU marching cubes over a distance volume (numpy) -- TEST INFRASTRUCTURE for the GPU mesher (csrc/mesh.cu).

The reference's mesher is the `isosurface` crate (un-vendored, /root/reference/Cargo.toml:91; call site
src/sdf/meshers/isosurface.rs:16-66): its vertex order and table cannot be restated here -- "parity unpinned" for
those.  What this file pins instead: (1) `extract` applies the case table of sdf-viewer_b200/mc_table.py to every
cell with plain loops-free numpy and the vertex rule below, so the GPU kernels' scan / ownership / indexing logic is
checked against an independent implementation of the same specification; (2) `check_manifold` / `signed_volume`
check properties no table error survives: closed 2-manifold, consistent outward orientation, enclosed volume;
(3) `postproc` restates Mesh::postproc (src/sdf/meshers/mesh.rs:22-33) and SDFSurface::normal's default
(src/sdf/defaults.rs:49-56) on top of the oracle's `sample`.

Vertex rule: an edge between lattice points a (inside, d < 0) and b (outside), or b inside and a outside, with
a < b along the edge's axis, carries ONE vertex at pos(a) + t * (pos(b) - pos(a)), t = (0 - d_a) / (d_b - d_a),
all in f32, where d = tex0.r - 0.1 (material.frag:56-60).
"""
import numpy as np

import sdf_viewer_b200.mc_table as T

f32 = np.float32


def lattice_positions(bb, dims):
    """voxel positions per axis, scene/sdf/mod.rs:179-182 (three separately rounded f32 operations)."""
    out = []
    for a in range(3):
        i = np.arange(dims[a], dtype=f32)
        s = f32(dims[a]) - f32(1.0)
        out.append(((i / s) * (f32(bb[1][a]) - f32(bb[0][a])) + f32(bb[0][a])).astype(f32))
    return out


def extract(tex0_r, bb, dims):
    """tex0_r: (D, H, W) float32 stored distances.  Returns (positions (n, 3) f32, triangles (m, 3) int64) with the
    vertices ordered by (owner lattice point in flat z-major order, axis) and the triangles by (cell, table order)."""
    W, H, D = dims
    d = (tex0_r.astype(f32) - f32(0.1)).astype(f32)
    inside = d < 0
    px, py, pz = lattice_positions(bb, dims)
    vid = np.full((D, H, W, 3), -1, np.int64)
    keys, pos = [], []
    for axis in range(3):
        sl_a = [slice(None)] * 3
        sl_b = [slice(None)] * 3
        np_axis = 2 - axis  # arrays are (z, y, x)
        sl_a[np_axis] = slice(0, -1)
        sl_b[np_axis] = slice(1, None)
        cross = inside[tuple(sl_a)] != inside[tuple(sl_b)]
        z, y, x = np.nonzero(cross)
        da = d[tuple(sl_a)][z, y, x]
        db = d[tuple(sl_b)][z, y, x]
        t = ((f32(0.0) - da) / (db - da)).astype(f32)
        p = np.stack([px[x], py[y], pz[z]], 1).astype(f32)
        tab = (px, py, pz)[axis]
        idx = (x, y, z)[axis]
        p[:, axis] = (tab[idx] + t * (tab[idx + 1] - tab[idx])).astype(f32)
        flat = (z.astype(np.int64) * H + y) * W + x
        keys.append(flat * 3 + axis)
        pos.append(p)
    keys = np.concatenate(keys)
    pos = np.concatenate(pos)
    order = np.argsort(keys, kind="stable")
    keys, pos = keys[order], pos[order]
    vid.reshape(-1)[keys] = np.arange(len(keys))
    # cells
    case = np.zeros((D - 1, H - 1, W - 1), np.int64)
    for c in range(8):
        dx, dy, dz = c & 1, (c >> 1) & 1, (c >> 2) & 1
        case |= inside[dz:D - 1 + dz, dy:H - 1 + dy, dx:W - 1 + dx].astype(np.int64) << c
    count = np.array([len(t) for t in T.TRI_TABLE])
    cz, cy, cx = np.nonzero(count[case] > 0)
    cc = case[cz, cy, cx]
    tris = []
    cell_order = []
    for k in range(T.MAX_TRIS):
        sel = count[cc] > k
        if not sel.any():
            break
        tri = np.empty((int(sel.sum()), 3), np.int64)
        for j in range(3):
            e = np.array([t[k][j] if len(t) > k else 0 for t in T.TRI_TABLE])[cc[sel]]
            own = np.array(T.EDGE_OWNER)[e]
            tri[:, j] = vid[cz[sel] + own[:, 2], cy[sel] + own[:, 1], cx[sel] + own[:, 0], own[:, 3]]
        tris.append(tri)
        cell_order.append(((cz[sel] * (H - 1) + cy[sel]) * (W - 1) + cx[sel]) * T.MAX_TRIS + k)
    if tris:
        tris = np.concatenate(tris)
        tris = tris[np.argsort(np.concatenate(cell_order), kind="stable")]
    else:
        tris = np.zeros((0, 3), np.int64)
    assert (tris >= 0).all()
    return pos, tris


def check_manifold(tris, closed=True):
    """Every directed edge appears once, and (closed) its reverse appears once: a consistently oriented 2-manifold
    without boundary.  Returns the number of boundary edges (0 when closed)."""
    a = np.concatenate([tris[:, [0, 1]], tris[:, [1, 2]], tris[:, [2, 0]]])
    nv = int(tris.max()) + 1 if len(tris) else 0
    fwd = a[:, 0] * nv + a[:, 1]
    rev = a[:, 1] * nv + a[:, 0]
    assert len(np.unique(fwd)) == len(fwd), "a directed edge is used by two triangles (inconsistent orientation or a fin)"
    boundary = int((~np.isin(fwd, rev)).sum())
    if closed:
        assert boundary == 0, f"{boundary} boundary edges: the mesh is not watertight"
    return boundary


def signed_volume(pos, tris):
    p = pos.astype(np.float64)
    a, b, c = p[tris[:, 0]], p[tris[:, 1]], p[tris[:, 2]]
    return float(np.einsum("ij,ij->i", a, np.cross(b, c)).sum() / 6.0)


def euler_characteristic(tris):
    nv = len(np.unique(tris))
    e = np.concatenate([tris[:, [0, 1]], tris[:, [1, 2]], tris[:, [2, 0]]])
    e.sort(axis=1)
    ne = len(np.unique(e[:, 0] * (int(tris.max()) + 1) + e[:, 1]))
    return nv - ne + len(tris)


def canonical_triangles(pos, tris):
    """Triangles as sorted rows of 9 position words (rotation-normalised): comparable across vertex orders."""
    p = pos.view(np.uint32).astype(np.int64)
    key = (p[:, 0] << 42) ^ (p[:, 1] << 21) ^ p[:, 2]
    t = tris.copy()
    k = key[t]
    first = np.argmin(k, axis=1)
    rows = np.arange(len(t))
    t = np.stack([t[rows, first], t[rows, (first + 1) % 3], t[rows, (first + 2) % 3]], 1)
    out = p[t].reshape(len(t), 9)
    return out[np.lexsort(out.T[::-1])]


def normal_default(sample_distance, p, eps=0.001):
    """SDFSurface::normal's default, src/sdf/defaults.rs:49-56: four tetrahedral taps, cgmath normalize
    (v * (1 / |v|)); f32 throughout, the sum in source order."""
    p = p.astype(f32)
    e = f32(eps)
    ks = np.array([[1, -1, -1], [-1, 1, -1], [-1, -1, 1], [1, 1, 1]], f32)
    acc = np.zeros_like(p)
    for k in ks:
        dk = sample_distance((p + k * e).astype(f32)).astype(f32)
        acc = (acc + k[None, :] * dk[:, None]).astype(f32)
    mag = np.sqrt(((acc[:, 0] * acc[:, 0] + acc[:, 1] * acc[:, 1]).astype(f32) + acc[:, 2] * acc[:, 2]).astype(f32)).astype(f32)
    return (acc * (f32(1.0) / mag)[:, None]).astype(f32)


def postproc(sample, positions):
    """Mesh::postproc, mesh.rs:22-33: (n, 12) vertex records -- position, normal, colour, metallic, roughness,
    occlusion -- from `sample(points) -> (n, 7)` (distance, r, g, b, metallic, roughness, occlusion)."""
    s = sample(positions)
    n = normal_default(lambda q: sample(q)[:, 0], positions)
    return np.concatenate([positions.astype(f32), n, s[:, 1:7].astype(f32)], 1)


PLY_HEADER = """ply
format ascii 1.0
comment {comment}
element vertex {nv}
property float x
property float y
property float z
property float nx
property float ny
property float nz
property uchar red
property uchar green
property uchar blue
property float metallic
property float roughness
property float occlusion
element face {nf}
property list uchar int vertex_index
end_header
"""
