"""GPU parity of the grid fill: the CUDA path through the C ABI against the CPU oracle,
bit for bit (f32 compared as u32), on the same tape / grid / bounding box.

Reference semantics: SDFViewer::update, /root/reference/src/app/scene/sdf/mod.rs:128-217."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

BB = ((-1.0, -1.0, -1.0), (1.0, 1.0, 1.0))


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def assert_same_volume(got0, got1, want0, want1):
    for name, g, w in (("tex0", got0, want0), ("tex1", got1, want1)):
        gb, wb = bits(g), bits(w)
        if not np.array_equal(gb, wb):
            bad = np.argwhere(gb != wb)
            z, y, x, c = bad[0]
            raise AssertionError(
                f"{name}: {len(bad)} of {gb.size} lanes differ; first at voxel ({x},{y},{z}) lane {c}: "
                f"got {g[z, y, x, c]!r} want {w[z, y, x, c]!r}")


def oracle_full(oracle, tape, dims, bb=BB, passes=2):
    v = oracle.Viewer(bb, dims, passes)
    s = oracle.Sampler(tape=tape)
    v.update(s)  # until the LoadingManager is exhausted, reference visit order
    return v


PROGRAMS = {"interpreter": (1, 0), "builtin_demo": (2, 2), "jit": (3, 1)}  # option value, last_fill_program


@pytest.mark.parametrize("program", list(PROGRAMS))
@pytest.mark.parametrize("vpt", [1, 2, 4, 8])
def test_demo_fill_all_64(S, oracle, vpt, program):
    """All three ways of executing the lowered tape (interpreter, built-in demo program, NVRTC
    kernel specialised for the tape structure) give the oracle's bits."""
    tape = S.tape.demo_tape()
    with S.SDFViewer.from_bb(BB, 64, 2) as v:
        assert v.dims == (64, 64, 64)
        v.set_option("fill_voxels_per_thread", vpt)
        v.set_option("fill_program", PROGRAMS[program][0])
        v.set_tape(tape)
        v.fill_all()
        assert v.get_info("last_fill_program") == PROGRAMS[program][1]
        assert v.get_info("last_fill_voxels_per_thread") == vpt
        t0, t1 = v.download()
        assert len(v.loading_mgr) == 0 and v.loading_mgr.passes_left() == 0
    o = oracle_full(oracle, tape, (64, 64, 64))
    assert_same_volume(t0, t1, o.tex0, o.tex1)
    # SURVEY 8c known answers (demo defaults, N = 64)
    assert t0[0, 0, 0, 0] == np.float32(0.1) + (np.float32(1.0) - np.float32(0.95))
    assert tuple(t1[0, 0, 0]) == (np.float32(0.4), np.float32(0.5), np.float32(1.0), np.float32(oracle.lib().orc_air_dist()))
    assert t0[32, 32, 32, 0] == np.float32(1.0)


@pytest.mark.parametrize("dims", [(2, 2, 2), (8, 8, 8), (11, 11, 11), (8, 11, 17), (33, 7, 5), (70, 9, 3)])
@pytest.mark.parametrize("passes", [1, 3])
def test_demo_update_passes_ragged(S, oracle, dims, passes):
    """Pass-by-pass: after every LoadingManager pass the volume equals the oracle's after the same
    number of iterations (loading.rs pass structure), including untouched AIR_DIST voxels."""
    tape = S.tape.demo_tape()
    o = oracle.Viewer(BB, dims, passes)
    s = oracle.Sampler(tape=tape)
    with S.SDFViewer.new_voxels(dims, BB, passes) as v:
        v.set_tape(tape)
        for step in S.loading.pass_steps(passes):
            n = S.loading.pass_items(dims, step)
            assert v.update(None, max_passes=1) == n
            assert o.update(s, max_iterations=n) == n
            t0, t1 = v.download()
            assert_same_volume(t0, t1, o.tex0, o.tex1)
            assert len(v.loading_mgr) == o.len()
            assert v.loading_mgr.total_iterations() == o.total_iterations()
            assert v.loading_mgr.passes_left() == o.passes_left()
        assert v.update(None) == 0 and o.update(s) == 0


def test_demo_params_and_disable_sphere(S, oracle):
    for kw in (dict(cube_half_side=0.8, sphere_radius=0.9), dict(disable_sphere=True),
               dict(cube_material=S.tape.MAT_NORMAL, sphere_material=S.tape.MAT_BRICK),
               dict(max_distance_custom_material=0.2)):
        tape = S.tape.demo_tape(**kw)
        with S.SDFViewer.from_bb(BB, 48, 2) as v:
            v.set_tape(tape)
            v.fill_all()
            t0, t1 = v.download()
        o = oracle_full(oracle, tape, (48, 48, 48))
        assert_same_volume(t0, t1, o.tex0, o.tex1)


def test_nonuniform_bbox(S, oracle):
    bb = ((-0.3, -1.1, 0.2), (1.7, 0.4, 0.9))
    dims = S.dims_from_bb(bb, 50)
    import ctypes as C
    od = (C.c_uint32 * 3)()
    oracle.lib().orc_dims_from_bb(oracle._bb6(bb), 50, od)
    assert dims == tuple(od)
    tape = S.tape.demo_tape()
    with S.SDFViewer.from_bb(bb, 50, 2) as v:
        assert v.dims == dims
        v.set_tape(tape)
        assert v.update(None) == sum(S.loading.pass_items(dims, s) for s in S.loading.pass_steps(2))
        t0, t1 = v.download()
    o = oracle_full(oracle, tape, dims, bb=bb)
    assert_same_volume(t0, t1, o.tex0, o.tex1)


@pytest.mark.parametrize("program", ["interpreter", "jit"])
@pytest.mark.parametrize("n_prims,vpt", [(5, 2), (40, 1), (40, 4), (300, 8), (1000, 8)])
def test_csg_tape(S, oracle, n_prims, vpt, program):
    """UNION_RANGE with per-tile culling (>= 16 primitives) and without: identical to the oracle's
    plain left-to-right fold."""
    table = S.tape.csg_primitive_table(n_prims, seed=7 + n_prims)
    tape = S.tape.csg_tape(table)
    dims = (64, 64, 32) if n_prims >= 300 else (40, 36, 24)
    with S.SDFViewer.new_voxels(dims, BB, 1) as v:
        v.set_option("fill_voxels_per_thread", vpt)
        v.set_option("fill_program", PROGRAMS[program][0])
        v.set_tape(tape)
        assert v.get_info("tape_culled") == (1 if n_prims >= 16 else 0)
        v.fill_all()
        assert v.get_info("last_fill_program") == PROGRAMS[program][1]
        t0, t1 = v.download()
        # sdfgpu_cull_stats: one more fill of every voxel (same volume), survivors per tile within [1, n_prims]
        st = v.cull_stats()
        vz = v.get_info("last_fill_voxels_per_thread")
        if n_prims >= 16:
            assert st["primitives"] == n_prims and 1 <= st["survivors_mean"] <= st["survivors_max"] <= n_prims
            assert st["tiles"] == -(-dims[0] // 32) * -(-dims[1] // 8) * -(-dims[2] // vz)
        else:
            assert st == {"primitives": 0, "tiles": 0, "survivors_mean": 0.0, "survivors_max": 0}
        u0, u1 = v.download()
        assert_same_volume(u0, u1, t0, t1)
    o = oracle.Viewer(BB, dims, 1)
    o.fill_all(oracle.Sampler(tape=tape))
    assert_same_volume(t0, t1, o.tex0, o.tex1)


@pytest.mark.parametrize("program,vpt", [("jit", 8), ("jit", 2), ("interpreter", 4)])
def test_csg_coarse_cell_precull(S, oracle, program, vpt):
    """A culled UNION_RANGE on a grid of several 64^3-voxel cells (ragged in every axis): tiles start from their cell's
    survivors (option fill_cull_cells, default) -- the same volume as culling the whole range per tile, and as the
    oracle's plain fold; also through the two-pass progressive load (step-2 lattice tiles straddle cells) and a dirty
    box that starts off the tile grid."""
    n_prims = 300
    table = S.tape.csg_primitive_table(n_prims, seed=11)
    tape = S.tape.csg_tape(table)
    dims = (160, 130, 70)
    o = oracle.Viewer(BB, dims, 2)
    o.update(oracle.Sampler(tape=tape))
    vols = []
    for cells in (1, 0):
        with S.SDFViewer.new_voxels(dims, BB, 2) as v:
            v.set_option("fill_voxels_per_thread", vpt)
            v.set_option("fill_program", PROGRAMS[program][0])
            v.set_option("fill_cull_cells", cells)
            v.set_tape(tape)
            v.update(None)                       # pass of step 2, then step 1
            t0, t1 = v.download()
            assert_same_volume(t0, t1, o.tex0, o.tex1)
            st = v.cull_stats()                  # (one more fill_all)
            other = S.tape.csg_tape(S.tape.csg_primitive_table(40, seed=5))
            v.set_tape(other)
            v.resample_box((-0.37, -0.41, -0.29, 0.53, 0.22, 0.61))
            vols.append((v.download(), st))
    (a0, a1), st_cells = vols[0]
    (b0, b1), st_full = vols[1]
    assert_same_volume(a0, a1, b0, b1)
    # the primitive with the least upper bound over a tile is never dropped by its cell (bounds over a larger box are
    # wider), so the tile's threshold -- and with it the tile's survivors -- are the same with and without the pre-cull
    assert st_cells == st_full and st_cells["survivors_mean"] >= 1


@pytest.mark.parametrize("program", ["interpreter", "jit"])
def test_generic_ops_tape(S, oracle, program):
    """Every opcode of include/sdfgpu_tape.h at least once."""
    T = S.tape
    t = T.TapeBuilder()
    a = t.prim(T.SHAPE_SPHERE, (0.2, 0.1, -0.3), 0.5, T.MAT_NORMAL, air_skip=0.3)
    b = t.prim(T.SHAPE_BOX_LINF, (-0.2, 0.0, 0.1), 0.4, T.MAT_BRICK, air_skip=float("inf"))
    c = t.prim(T.SHAPE_BOX_LINF, (0.0, 0.0, 0.0), 0.9, T.MAT_FLAT, color=(0.2, 0.9, 0.1), metallic=0.3,
               roughness=0.7, occlusion=0.5)
    d = t.prim(T.SHAPE_SPHERE, (0.5, 0.5, 0.5), 0.3, T.MAT_FLAT, color=(1.5, -0.2, 0.0), occlusion=-1.0)
    k0 = t.const([0.1, -0.05, 0.02])
    k1 = t.const([0.9, 0.8, 0.7, 0.6, 0.5, 0.4])
    (t.emit(T.OP_PRIM, a).emit(T.OP_UNION_PRIM, b).emit(T.OP_PUSH)
      .emit(T.OP_P_SUB, k0).emit(T.OP_P_MUL, imm=1.5).emit(T.OP_P_ABS, 5)
      .emit(T.OP_PRIM, d).emit(T.OP_D_MUL, imm=0.5).emit(T.OP_D_ADD, imm=-0.01).emit(T.OP_POP_UNION)
      .emit(T.OP_PUSH).emit(T.OP_P_RESET).emit(T.OP_PRIM, c).emit(T.OP_D_NEG).emit(T.OP_D_ABS)
      .emit(T.OP_D_MAX, imm=-0.2).emit(T.OP_D_MIN, imm=0.7).emit(T.OP_POP_INTER)
      .emit(T.OP_PUSH).emit(T.OP_UNION_RANGE, 0, 4).emit(T.OP_INTER_PRIM, c).emit(T.OP_M_SET, k1)
      .emit(T.OP_POP_UNION).emit(T.OP_END))
    tape = t.build()
    dims = (37, 29, 13)
    with S.SDFViewer.new_voxels(dims, BB, 1) as v:
        v.set_option("fill_program", PROGRAMS[program][0])
        v.set_tape(tape)
        v.fill_all()
        assert v.get_info("last_fill_program") == PROGRAMS[program][1]
        t0, t1 = v.download()
    o = oracle.Viewer(BB, dims, 1)
    o.fill_all(oracle.Sampler(tape=tape))
    assert_same_volume(t0, t1, o.tex0, o.tex1)


@pytest.mark.parametrize("program", ["interpreter", "jit"])
def test_deep_stack_tape(S, oracle, program):
    """Stack depth 8 (SDFT_MAX_STACK): levels below the top are spilled to shared memory."""
    T = S.tape
    t = T.TapeBuilder()
    rng = np.random.default_rng(3)
    prims = [t.prim(int(rng.integers(0, 2)), rng.uniform(-0.6, 0.6, 3), float(rng.uniform(0.2, 0.5)),
                    int(rng.integers(0, 3)), color=rng.uniform(0, 1, 3), metallic=0.3, roughness=0.2, occlusion=0.9,
                    air_skip=float(rng.choice([0.1, np.inf]))) for _ in range(9)]
    for k in range(8):
        t.emit(T.OP_PRIM, prims[k]).emit(T.OP_PUSH)
    t.emit(T.OP_PRIM, prims[8])
    for k in range(8):
        t.emit(T.OP_POP_UNION if k % 2 == 0 else T.OP_POP_INTER)
    t.emit(T.OP_END)
    tape = t.build()
    dims = (33, 17, 9)
    for vpt in (1, 4):
        with S.SDFViewer.new_voxels(dims, BB, 1) as v:
            v.set_option("fill_program", PROGRAMS[program][0])
            v.set_option("fill_voxels_per_thread", vpt)
            v.set_tape(tape)
            v.fill_all()
            t0, t1 = v.download()
        o = oracle.Viewer(BB, dims, 1)
        o.fill_all(oracle.Sampler(tape=tape))
        assert_same_volume(t0, t1, o.tex0, o.tex1)


def random_tape(T, rng):
    """A random valid tape: random primitives, every opcode class, balanced stack (depth <= 8)."""
    t = T.TapeBuilder()
    n_prims = int(rng.integers(1, 40))
    prims = [t.prim(int(rng.integers(0, 2)), rng.uniform(-0.9, 0.9, 3), float(rng.uniform(0.05, 0.8)),
                    int(rng.integers(0, 3)), color=rng.uniform(-0.2, 1.3, 3), metallic=float(rng.uniform(0, 1)),
                    roughness=float(rng.uniform(0, 1)), occlusion=float(rng.uniform(-0.5, 1)),
                    air_skip=float(rng.choice([0.05, 0.1, 0.5, np.inf]))) for _ in range(n_prims)]
    consts = t.const(rng.uniform(-0.5, 0.9, 16))
    depth = 0
    t.emit(T.OP_PRIM, int(rng.integers(0, n_prims)))
    for _ in range(int(rng.integers(3, 40))):
        r = rng.random()
        if r < 0.25:
            t.emit(int(rng.choice([T.OP_PRIM, T.OP_UNION_PRIM, T.OP_INTER_PRIM])), int(rng.integers(0, n_prims)))
        elif r < 0.32:
            a = int(rng.integers(0, n_prims)); b = int(rng.integers(1, n_prims - a + 1))
            t.emit(T.OP_UNION_RANGE, a, b)
        elif r < 0.47 and depth < 8:
            t.emit(T.OP_PUSH); depth += 1
            t.emit(T.OP_PRIM, int(rng.integers(0, n_prims)))
        elif r < 0.62 and depth > 0:
            op = int(rng.choice([T.OP_POP_UNION, T.OP_POP_INTER, T.OP_POP_DEMO_DIFF]))
            t.emit(op, int(rng.integers(0, 9)) if op == T.OP_POP_DEMO_DIFF else 0); depth -= 1
        elif r < 0.77:
            op = int(rng.choice([T.OP_D_NEG, T.OP_D_ABS, T.OP_D_ADD, T.OP_D_MUL, T.OP_D_MAX, T.OP_D_MIN]))
            t.emit(op, imm=float(rng.uniform(-0.5, 1.5)))
        elif r < 0.83:
            t.emit(T.OP_M_SET, int(rng.integers(0, 10)))
        else:
            op = int(rng.choice([T.OP_P_RESET, T.OP_P_SUB, T.OP_P_MUL, T.OP_P_ABS]))
            t.emit(op, int(rng.integers(0, 8)) if op == T.OP_P_ABS else int(rng.integers(0, 13)), imm=float(rng.uniform(0.5, 2.0)))
    while depth > 0:
        t.emit(int(rng.choice([T.OP_POP_UNION, T.OP_POP_INTER]))); depth -= 1
    if rng.random() < 0.7:
        t.emit(T.OP_END)
    return t.build()


@pytest.mark.parametrize("seed", range(12))
def test_random_tapes(S, oracle, seed):
    """Fuzz: random tapes on random ragged grids and boxes; interpreter and specialised kernel both equal
    the oracle's interpreter bit for bit."""
    rng = np.random.default_rng(1000 + seed)
    tape = random_tape(S.tape, rng)
    dims = tuple(int(x) for x in rng.integers(1, 40, 3))
    lo = rng.uniform(-1.5, -0.2, 3); hi = lo + rng.uniform(0.3, 3.0, 3)
    bb = (tuple(float(x) for x in lo), tuple(float(x) for x in hi))
    o = oracle.Viewer(bb, dims, 1)
    with np.errstate(all="ignore"):
        o.fill_all(oracle.Sampler(tape=tape))
    for program in ("interpreter", "jit"):
        with S.SDFViewer.new_voxels(dims, bb, 1) as v:
            v.set_option("fill_program", PROGRAMS[program][0])
            v.set_option("fill_voxels_per_thread", int(rng.choice([0, 1, 2, 4, 8])))
            v.set_tape(tape)
            v.fill_all()
            t0, t1 = v.download()
        ok0 = (bits(t0) == bits(o.tex0)) | (np.isnan(t0) & np.isnan(o.tex0))
        ok1 = (bits(t1) == bits(o.tex1)) | (np.isnan(t1) & np.isnan(o.tex1))
        assert ok0.all() and ok1.all(), (program, seed, int((~ok0).sum()), int((~ok1).sum()))


def test_changed_box_state_machine(S, oracle):
    """sdf.changed() -> merged pending box -> 3-pass re-sample of the voxels inside it
    (scene/sdf/mod.rs:131-154,184-190), while and after loading.  The GPU runs whole passes; the
    oracle is then given the same number of iterations, so both stop at the same pass boundary."""
    import ctypes as C
    dims = (24, 20, 16)
    sdf = S.SDFDemo()
    o = oracle.Viewer(BB, dims, 2)
    with S.SDFViewer.new_voxels(dims, BB, 2) as v:
        def step(changed, tape, max_passes):
            v.set_tape(tape)
            it = C.c_uint64()
            box = (C.c_float * 6)(*changed) if changed is not None else None
            S.viewer.check(v._lib.sdfgpu_update(v._h, box, max_passes, C.byref(it)), v._h)
            s = oracle.Sampler(tape=tape)
            want = o.update(s, changed=changed, max_iterations=it.value if it.value else 1)
            assert it.value == want
            t0, t1 = v.download()
            assert_same_volume(t0, t1, o.tex0, o.tex1)
            assert len(v.loading_mgr) == o.len()
            assert v.loading_mgr.passes_left() == o.passes_left()
            return it.value

        tapeA = sdf.tape()
        assert step(None, tapeA, 1) == S.loading.pass_items(dims, 2)   # coarse pass of the initial load
        boxA = (-0.5, -0.25, -1.0, 0.25, 0.5, 0.1)
        sdf.set_parameter("sphere_radius", 0.9)
        tapeB = sdf.tape()
        assert step(boxA, tapeB, 1) > 0            # change reported while loading
        assert step(None, tapeB, 0) > 0            # queued 3-pass re-sample (changed_box_while_loading)
        assert step(None, tapeB, 0) > 0            # one more 3-pass round, then the box is dropped
        assert step(None, tapeB, 0) == 0
        boxB = (0.1, 0.1, 0.1, 0.9, 0.6, 0.7)
        sdf.set_parameter("cube_half_side", 0.7)
        tapeC = sdf.tape()
        assert step(boxB, tapeC, 2) > 0            # change after loading: new 3-pass manager
        assert step(boxA, tapeC, 0) > 0            # merged boxes
        step(None, tapeC, 0)
        step(None, tapeC, 0)
        assert step(None, tapeC, 0) == 0
        assert len(v.loading_mgr) == 0


def test_resample_box(S, oracle):
    """The dirty-block path: only voxels whose position lies in the closed box change."""
    dims = (48, 40, 32)
    tapeA, tapeB = S.tape.demo_tape(), S.tape.demo_tape(sphere_radius=0.8, cube_half_side=0.9)
    box = (-0.4, -0.35, -0.6, 0.31, 0.55, 0.2)
    with S.SDFViewer.new_voxels(dims, BB, 2) as v:
        v.set_tape(tapeA)
        v.fill_all()
        v.set_tape(tapeB)
        n = v.resample_box(box, count=True)
        t0, t1 = v.download()
    a = oracle.Viewer(BB, dims, 1); a.fill_all(oracle.Sampler(tape=tapeA))
    b = oracle.Viewer(BB, dims, 1); b.fill_all(oracle.Sampler(tape=tapeB))
    xs = np.array([a.voxel_pos(i, 0, 0)[0] for i in range(dims[0])], np.float32)
    ys = np.array([a.voxel_pos(0, i, 0)[1] for i in range(dims[1])], np.float32)
    zs = np.array([a.voxel_pos(0, 0, i)[2] for i in range(dims[2])], np.float32)
    f = np.float32
    inside = ((zs >= f(box[2])) & (zs <= f(box[5])))[:, None, None] & ((ys >= f(box[1])) & (ys <= f(box[4])))[None, :, None] \
        & ((xs >= f(box[0])) & (xs <= f(box[3])))[None, None, :]
    want0 = np.where(inside[..., None], b.tex0, a.tex0)
    want1 = np.where(inside[..., None], b.tex1, a.tex1)
    assert_same_volume(t0, t1, want0, want1)
    assert n >= inside.sum()  # AIR_DIST-valued voxels inside the index box are re-sampled too


def test_slab_handles_cover_grid(S, oracle):
    """Z-slab handles (multi-GPU sharding unit) each produce exactly their slices of the full grid,
    and locally computed halo slices equal the neighbour's owned slices bit for bit."""
    dims = (32, 24, 20)
    tape = S.tape.demo_tape()
    o = oracle_full(oracle, tape, dims)
    cuts = [0, 7, 8, 15, 20]
    for zb, ze in zip(cuts[:-1], cuts[1:]):
        with S.SDFViewer.new_voxels(dims, BB, 2, z_range=(zb, ze)) as v:
            assert (v.z_begin, v.z_end) == (zb, ze)
            assert v.z_lo == max(zb - 1, 0) and v.z_hi == min(ze + 1, dims[2])
            v.set_tape(tape)
            v.update(None)
            t0, t1 = v.download()
            assert_same_volume(t0, t1, o.tex0[zb:ze], o.tex1[zb:ze])


def test_empty_and_degenerate_grids(S, oracle):
    """A bounding box with a zero-size axis gives 0 voxels on it (scene/sdf/mod.rs:54-64): every call
    still succeeds; a 1-voxel axis puts the voxel at 0/0 = NaN like the reference's position formula."""
    bb = ((-1.0, 0.0, -1.0), (1.0, 0.0, 1.0))
    assert S.dims_from_bb(bb, 32) == (32, 0, 32)
    with S.SDFViewer.from_bb(bb, 32, 2) as v:
        v.set_tape(S.tape.demo_tape())
        assert v.update(None) == 0
        v.fill_all()
        v.commit()
        t0, t1 = v.download()
        assert t0.size == 0 and t1.size == 0
        r, d, g = v.trace(S.default_camera(64, 48), 64, 48, gbuf=True)
        assert np.all(r == 0) and np.all(d == 1) and np.all(g[..., 3] == -3)
    dims = (5, 1, 4)
    with S.SDFViewer.new_voxels(dims, BB, 1) as v:
        v.set_tape(S.tape.demo_tape())
        v.fill_all()
        t0, t1 = v.download()
    o = oracle.Viewer(BB, dims, 1)
    with np.errstate(all="ignore"):
        o.fill_all(oracle.Sampler(tape=S.tape.demo_tape()))
    same = (bits(t0) == bits(o.tex0)) | (np.isnan(t0) & np.isnan(o.tex0))
    assert same.all() and np.array_equal(bits(t1), bits(o.tex1))


def test_batched_ingest_of_host_samples(S, oracle):
    """an SDF without a tape is sampled on the host in batches (here: the oracle's
    SDFDemo::sample standing in for a WASM guest) at the positions the library reports; the GPU applies
    the store rules.  The volume equals the oracle's; odd records (black, NaN, out-of-range colours,
    negative occlusion) follow scene/sdf/mod.rs:196-208 exactly."""
    dims = (21, 13, 9)
    n = dims[0] * dims[1] * dims[2]
    o = oracle.Viewer(BB, dims, 1)
    o.fill_all(oracle.Sampler())
    with S.SDFViewer.new_voxels(dims, BB, 1) as v:
        pos = v.voxel_positions(0, n)
        want_pos = np.array([o.voxel_pos(x, y, z) for z in range(dims[2]) for y in range(dims[1]) for x in range(dims[0])], np.float32)
        assert np.array_equal(bits(pos), bits(want_pos))
        for first in range(0, n, 1000):  # batches, as a host loop would send them
            chunk = pos[first:first + 1000]
            v.ingest_samples(first, oracle.demo_sample(chunk))
        t0, t1 = v.download()
        assert_same_volume(t0, t1, o.tex0, o.tex1)
        odd = np.array([[0.3, 0, 0, 0, 0.1, 0.2, 0.0], [np.nan, 1.5, -0.2, np.nan, 2.0, -1.0, -3.0],
                        [-5.0, 0.5, 0.6, 0.7, 0.5, 0.0, 0.25], [7.0, 1.0, 1.0, 1.0, 0, 0, 1e-30]], np.float32)
        v.ingest_samples(5, odd)
        t0, t1 = v.download()
    lut = (oracle.C.c_float * 256)(); oracle.lib().orc_srgb_lut(lut)
    f0, f1 = t0.reshape(-1, 4)[5:9], t1.reshape(-1, 4)[5:9]
    assert tuple(f0[0]) == (np.float32(0.1) + np.float32(0.3), lut[127], lut[127], lut[127]) and f1[0][2] == 1.0
    assert np.isnan(f0[1][0]) and tuple(f0[1][1:]) == (lut[255], lut[0], lut[0]) and tuple(f1[1][:3]) == (2.0, -1.0, 1.0)
    assert f0[2][0] == 0.0 and f0[3][0] == 1.0 and f1[2][2] == np.float32(0.25) and f1[3][2] == np.float32(1e-30)
    with S.SDFViewer.new_voxels(dims, BB, 1, z_range=(3, 6)) as v:
        with pytest.raises(S.SdfGpuError):
            v.ingest_samples(0, odd)  # slice 0 is not stored by this slab


def test_known_state_tracking_mixed_calls(S, oracle):
    """The read-free passes rely on what the host knows about the volume; calls that break that
    knowledge (ingest / resample in the middle of a load) must fall back to reading tex0.r.  Expected
    volume built by hand from the rule of scene/sdf/mod.rs:184-190."""
    dims = (24, 16, 12)
    n = dims[0] * dims[1] * dims[2]
    tapeA, tapeB = S.tape.demo_tape(), S.tape.demo_tape(cube_half_side=0.6, sphere_radius=0.7)
    a = oracle.Viewer(BB, dims, 1); a.fill_all(oracle.Sampler(tape=tapeA))
    b = oracle.Viewer(BB, dims, 1); b.fill_all(oracle.Sampler(tape=tapeB))
    air = np.float32(oracle.lib().orc_air_dist())
    with S.SDFViewer.new_voxels(dims, BB, 2) as v:
        v.set_tape(tapeA)
        assert v.update(None, max_passes=1) == S.loading.pass_items(dims, 2)      # coarse lattice, no read
        pos = v.voxel_positions(100, 700)
        v.ingest_samples(100, oracle.tape_sample(tapeB, pos))                      # knowledge lost
        assert v.update(None) == n                                                # fine pass must read
        t0, t1 = v.download()
        want0, want1 = a.tex0.copy().reshape(-1, 4), a.tex1.copy().reshape(-1, 4)
        want0[100:800], want1[100:800] = b.tex0.reshape(-1, 4)[100:800], b.tex1.reshape(-1, 4)[100:800]
        # the coarse lattice was sampled with tape A before the ingest overwrote part of it; ingested voxels
        # whose stored distance happens to equal AIR_DIST would be re-sampled (none here)
        assert not np.any(b.tex0.reshape(-1, 4)[100:800, 0] == air)
        assert_same_volume(t0, t1, want0.reshape(t0.shape), want1.reshape(t1.shape))
        # now fully sampled: a box pass touches only the box, a plain pass nothing
        v.set_tape(tapeB)
        assert v.resample_box((-0.3, -0.3, -0.3, 0.3, 0.3, 0.3), count=True) > 0
        assert v.update(None) == 0
        v.reset(2)                                                                # all AIR_DIST again
        assert v.update(None) == S.loading.pass_items(dims, 2) + n
        t0, t1 = v.download()
        assert_same_volume(t0, t1, b.tex0, b.tex1)


def test_errors(S):
    import ctypes as C
    lib = S.viewer._lib.load()
    with S.SDFViewer.from_bb(BB, 16, 2) as v:
        with pytest.raises(S.SdfGpuError) as e:
            v.fill_all()
        assert e.value.code == -4  # SDFGPU_ERR_STATE: no tape
        with pytest.raises(S.SdfGpuError) as e:
            v.set_tape(b"\0" * 64)
        assert e.value.code == -3
        bad = bytearray(S.tape.demo_tape())
        bad[32 + 4] = 99  # instr 0 operand a -> primitive out of range
        with pytest.raises(S.SdfGpuError) as e:
            v.set_tape(bytes(bad))
        assert e.value.code == -3 and "out of range" in e.value.message
        t = S.tape.TapeBuilder()
        t.emit(S.tape.OP_POP_UNION)
        with pytest.raises(S.SdfGpuError):
            v.set_tape(t.build())
        with pytest.raises(S.SdfGpuError):
            v.set_option("nope", 1)
    h = C.c_void_p()
    assert lib.sdfgpu_create((C.c_float * 6)(-1, -1, -1, 1, 1, 1), 16, 2, 999, C.byref(h)) == -1


def test_jit_cache_is_bounded():
    """The cache of specialised kernels keeps the most recently used ones and unloads the rest (jit.cu): run with a
    bound of 2 over 4 tape structures, twice (tests/jit_cache_check.py; the bound is read once per process)."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "tests", "jit_cache_check.py")], capture_output=True, text=True,
                       timeout=600, env=dict(os.environ, SDFGPU_JIT_CACHE="2"))
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "jit_cache_check ok" in r.stdout
