"""The mesher (reference: src/sdf/meshers/isosurface.rs:16-99, mesh.rs:22-33,38-129).

CPU: the derived marching-cubes table's invariants, the numpy restatement (tests/mc_ref.py) on analytic volumes,
and the PLY writer against the format mesh.rs defines.  GPU: sdfgpu_mesh against mc_ref.extract on the same volume
(same vertex set and triangle set, bit for bit), watertightness / orientation / volume, and the per-vertex records
against the oracle's sample() and the restated default normal.

Parity statement: the `isosurface` crate and `ply-rs` are not vendored with the reference (Cargo.toml:91,
Cargo.lock "ply-rs 0.1.3"), and the reference ships no mesh fixture: vertex order, the triangulation of ambiguous
cells and the byte-level PLY layout are "parity unpinned"; positions, materials and normals follow the reference's
formulas and are checked against their restatement."""
import os

import numpy as np
import pytest

BB = ((-1.0, -1.0, -1.0), (1.0, 1.0, 1.0))


def test_case_table_invariants(S):
    import sdf_viewer_b200.mc_table as T
    assert T.MAX_TRIS == 5 and T.TRI_TABLE[0] == [] and T.TRI_TABLE[255] == []
    for case in range(256):
        ins = [(case >> c) & 1 for c in range(8)]
        cross = set()
        for e in range(12):
            dx, dy, dz, axis = T.EDGE_OWNER[e]
            c0 = T.corner(dx, dy, dz)
            c1 = c0 | (1 << axis)
            if ins[c0] != ins[c1]:
                cross.add(e)
        used = [e for t in T.TRI_TABLE[case] for e in t]
        assert set(used) == cross, case                       # every sign-changing edge carries a vertex, no other does
        assert all(len(set(t)) == 3 for t in T.TRI_TABLE[case])
        # within the cell, every directed triangle edge that is not on a cube face is matched by its reverse
        directed = {}
        for t in T.TRI_TABLE[case]:
            for a, b in ((t[0], t[1]), (t[1], t[2]), (t[2], t[0])):
                assert (a, b) not in directed, case
                directed[(a, b)] = True
        def on_one_face(a, b):
            def faces_of(e):
                dx, dy, dz, axis = T.EDGE_OWNER[e]
                p = [dx, dy, dz]
                return {(k, p[k]) for k in range(3) if k != axis}
            return bool(faces_of(a) & faces_of(b))
        for (a, b) in directed:
            if (b, a) not in directed:
                assert on_one_face(a, b), (case, a, b)          # an open edge lies on a cube face (shared with the neighbour cell)
    # the generated include is current
    inc = open(T.INC).read()
    T.write_inc()
    assert open(T.INC).read() == inc


def test_reference_extraction_on_analytic_volumes(S):
    import mc_ref
    for dims, r in (((33, 33, 33), 0.7), ((40, 36, 50), 0.55), ((17, 64, 23), 0.8)):
        px, py, pz = mc_ref.lattice_positions(BB, dims)
        Z, Y, X = np.meshgrid(pz, py, px, indexing="ij")
        d = (np.sqrt(X * X + Y * Y + Z * Z) - np.float32(r)).astype(np.float32)
        tex = np.clip(d + np.float32(0.1), 0, 1).astype(np.float32)
        pos, tris = mc_ref.extract(tex, BB, dims)
        mc_ref.check_manifold(tris)
        assert mc_ref.euler_characteristic(tris) == 2
        vol = mc_ref.signed_volume(pos, tris)
        assert 0.97 < vol / (4 / 3 * np.pi * r ** 3) <= 1.0      # inscribed polyhedron, oriented outwards
        assert np.abs(np.linalg.norm(pos.astype(np.float64), axis=1) - r).max() < 0.6 * 2 / (min(dims) - 1)  # within half a cell diagonal
    # a sphere cut by the box: open mesh, boundary on the box faces only
    dims = (24, 24, 24)
    px, py, pz = mc_ref.lattice_positions(BB, dims)
    Z, Y, X = np.meshgrid(pz, py, px, indexing="ij")
    tex = np.clip(np.sqrt(X * X + Y * Y + Z * Z) - np.float32(1.2) + np.float32(0.1), 0, 1).astype(np.float32)
    pos, tris = mc_ref.extract(tex, BB, dims)
    assert mc_ref.check_manifold(tris, closed=False) > 0


def test_ply_writer(S, tmp_path):
    """Header and record layout of Mesh::serialize_ply (mesh.rs:45-121); floats as Rust's `{}` prints them."""
    import mc_ref
    v = np.zeros((3, 12), np.float32)
    v[0] = [0.5, -1.0, 1e-7, 0.0, -0.0, 1.0, 0.1, 0.999999, 1.5, 0.25, 0.0, 1.0]
    v[1] = [123456.79, 1 / 3, -2.5e-5, 0.57735026, 0.57735026, -0.57735026, 0.0, 0.5, 1.0, 0.0, 0.8, 0.3]
    v[2] = [3.4e38, 16777216.0, 1e10, np.nan, np.inf, -np.inf, -0.1, np.nan, 0.00390625, 1.0, 1.0, 1.0]
    t = np.array([[0, 1, 2], [2, 1, 0]], np.uint32)
    path = tmp_path / "m.ply"
    n = S.ply_serialize(v, t, path, comment="Created with test")
    text = open(path).read()
    assert n == len(text.encode())
    head = mc_ref.PLY_HEADER.format(comment="Created with test", nv=3, nf=2)
    assert text.startswith(head)
    body = text[len(head):].splitlines()
    assert body[0] == "0.5 -1 0.0000001 0 -0 1 25 255 255 0.25 0 1"
    assert body[1] == "123456.79 0.33333334 -0.000025 0.57735026 0.57735026 -0.57735026 0 127 255 0 0.8 0.3"
    assert body[2] == "340000000000000000000000000000000000000 16777216 10000000000 NaN inf -inf 0 0 0 1 1 1"
    assert body[3:] == ["3 0 1 2", "3 2 1 0"]
    # every finite float round-trips through its text
    rng = np.random.default_rng(5)
    r = rng.standard_normal((2000, 12)).astype(np.float32) * np.float32(10.0) ** rng.integers(-6, 6, (2000, 12)).astype(np.float32)
    S.ply_serialize(r, np.zeros((0, 3), np.uint32), path, comment="")
    lines = open(path).read().split("end_header\n")[1].splitlines()
    assert "comment" not in open(path).read().split("end_header")[0]
    back = np.array([[np.float32(x) for x in ln.split()] for ln in lines], np.float32)
    cols = [0, 1, 2, 3, 4, 5, 9, 10, 11]
    assert np.array_equal(back[:, cols].view(np.uint32), r[:, cols].view(np.uint32))
    assert not any("e" in ln for ln in lines)
    want_rgb = np.clip(np.nan_to_num(r[:, 6:9] * np.float32(255.9999)), 0, 255).astype(np.uint8)
    assert np.array_equal(back[:, 6:9].astype(np.uint8), want_rgb)


gpu = pytest.mark.gpu


@gpu
@pytest.mark.parametrize("dims", [(33, 33, 33), (40, 36, 50), (65, 65, 65), (70, 9, 130)])
def test_gpu_mesh_equals_reference_extraction(S, oracle, dims):
    import mc_ref
    sdf = S.SDFDemo()
    with S.SDFViewer.new_voxels(dims, BB, 1) as v:
        v.set_tape(sdf.tape()); v.fill_all(); v.commit()
        verts, tris = v.mesh()
        t0, _ = v.download()
    pos, want_tris = mc_ref.extract(np.ascontiguousarray(t0[..., 0]), BB, dims)
    assert len(verts) == len(pos) and len(tris) == len(want_tris)
    tris = tris.astype(np.int64)
    # the same vertex set and the same triangles (as position triples), bit for bit
    got_sorted = verts[:, :3][np.lexsort(verts[:, :3].view(np.uint32).T[::-1])]
    want_sorted = pos[np.lexsort(pos.view(np.uint32).T[::-1])]
    assert np.array_equal(got_sorted.view(np.uint32), want_sorted.view(np.uint32))
    assert np.array_equal(mc_ref.canonical_triangles(np.ascontiguousarray(verts[:, :3]), tris), mc_ref.canonical_triangles(pos, want_tris))
    # properties that do not depend on the table: closed, consistently oriented, outward
    if min(dims) > 16:
        mc_ref.check_manifold(tris)
        assert mc_ref.signed_volume(verts[:, :3], tris) > 1.5
        assert mc_ref.euler_characteristic(tris) == -8  # the demo: a cube with a sphere cut through its six faces (genus 5)
    # every vertex lies on a lattice edge whose ends have opposite signs, i.e. within one cell of the surface
    d = oracle.demo_sample(verts[:, :3])[:, 0]
    cell = max(2.0 / (n - 1) for n in dims)
    assert np.abs(d).max() < cell
    # vertex records: Mesh::postproc + SDFSurface::normal at the GPU's own positions
    want = mc_ref.postproc(lambda q: oracle.demo_sample(q), np.ascontiguousarray(verts[:, :3]))
    assert np.array_equal(verts[:, 6:12].view(np.uint32), want[:, 6:12].view(np.uint32)), "materials differ from sample(p, false)"
    ok = np.isfinite(want[:, 3:6]).all(axis=1)
    np.testing.assert_allclose(verts[ok, 3:6], want[ok, 3:6], rtol=0, atol=2e-6)
    assert (np.isfinite(verts[:, 3:6]).all(axis=1) == ok).all()
    # normals point out of the solid: they agree with the face normals of the triangles around them
    a, b, c = (verts[tris[:, k], :3].astype(np.float64) for k in range(3))
    fn = np.cross(b - a, c - a)
    big = np.linalg.norm(fn, axis=1) > 1e-9
    fn = fn[big] / np.linalg.norm(fn[big], axis=1)[:, None]
    vn = verts[:, 3:6][tris[big]].astype(np.float64).mean(axis=1)
    assert np.nanmean((fn * vn).sum(axis=1)) > 0.8


@gpu
def test_gpu_sample_points_is_sample(S, oracle):
    """sdfgpu_sample_points = SDFSurface::sample(p, false) for the demo (oracle demo_sample), a CSG tape (oracle tape
    evaluator; per-tile culling off in point mode) and a scalar-program tape."""
    rng = np.random.default_rng(11)
    pts = (rng.random((20011, 3), np.float32) * 2.4 - 1.2).astype(np.float32)
    with S.SDFViewer.new_voxels((16, 16, 16), BB, 1) as v:
        v.set_tape(S.tape.demo_tape())
        got = v.sample_points(pts)
        assert np.array_equal(got.view(np.uint32), oracle.demo_sample(pts).view(np.uint32))
        for prog in (1, 2):  # interpreter, built-in
            v.set_option("fill_program", prog)
            assert np.array_equal(v.sample_points(pts).view(np.uint32), got.view(np.uint32))
        v.set_option("fill_program", 0)
        csg = S.tape.csg_tape(S.tape.csg_primitive_table(200))
        v.set_tape(csg)
        assert v.get_info("tape_culled") == 1
        assert np.array_equal(v.sample_points(pts).view(np.uint32), oracle.tape_sample(csg, pts).view(np.uint32))
        assert v.sample_points(np.zeros((0, 3), np.float32)).shape == (0, 7)


@gpu
def test_gpu_mesh_ply_and_errors(S, tmp_path):
    import mc_ref
    with S.SDFViewer.new_voxels((33, 33, 33), BB, 1) as v:
        with pytest.raises(S.SdfGpuError):
            v.mesh()  # no tape
        v.set_tape(S.tape.demo_tape()); v.fill_all()
        with pytest.raises(S.SdfGpuError):
            v.mesh_write_ply(tmp_path / "x.ply")  # no mesh yet
        verts, tris = v.mesh()
        n = v.mesh_write_ply(tmp_path / "demo.ply", comment="Created with sdf-viewer_b200 test")
        text = open(tmp_path / "demo.ply").read()
        assert n == len(text)
        head = mc_ref.PLY_HEADER.format(comment="Created with sdf-viewer_b200 test", nv=len(verts), nf=len(tris))
        assert text.startswith(head)
        lines = text[len(head):].splitlines()
        assert len(lines) == len(verts) + len(tris)
        back = np.array([[np.float32(x) for x in ln.split()] for ln in lines[:len(verts)]], np.float32)
        assert np.array_equal(back[:, :6].view(np.uint32), verts[:, :6].view(np.uint32))
        faces = np.array([[int(x) for x in ln.split()] for ln in lines[len(verts):]])
        assert (faces[:, 0] == 3).all() and np.array_equal(faces[:, 1:], tris.astype(np.int64))
        # an empty volume (all air) gives an empty mesh
        v.reset(1)
        assert v.mesh(download=False) == (0, 0)
    with S.SDFViewer.new_voxels((16, 16, 16), BB, 1, z_range=(4, 9)) as slab:
        slab.set_tape(S.tape.demo_tape()); slab.fill_all()
        with pytest.raises(S.SdfGpuError):
            slab.mesh()
