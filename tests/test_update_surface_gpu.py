"""GPU parity of `sdfgpu_update_surface`: SDFViewer::update(sdf: impl SDFSurface, max_delta_time)
(/root/reference/src/app/scene/sdf/mod.rs:128-217) called with the trait object itself.

A surface with a tape runs on the GPU; a surface without one (any existing .wasm SDF) is sampled
on the host through its `sample` callback -- here the oracle's restatement of SDFDemo::sample stands
in for the WASM guest -- in the reference's LoadingManager order, chunk by chunk, and the results
are scattered into the volumes on the GPU.  Either way the volumes must equal the oracle's update
bit for bit and the loading counters must follow loading.rs."""
import numpy as np
import pytest

from test_fill_gpu import BB, assert_same_volume

pytestmark = pytest.mark.gpu


class HostDemo:
    """An SDFSurface WITHOUT a tape: sample() is host code (the oracle's SDFDemo::sample)."""

    def __init__(self, S, oracle, **params):
        self._S, self._oracle = S, oracle
        self.params = dict(params)
        self._changed = None
        self.calls = 0
        self.points = 0

    def bounding_box(self):
        return BB

    def _oracle_params(self):
        return self._oracle.demo_params(**self.params)

    def sample(self, points, distance_only=False):
        self.calls += 1
        self.points += len(points)
        return self._oracle.demo_sample(points, self._oracle_params(), distance_only)

    def set(self, box=None, **params):
        self.params.update(params)
        self._changed = box if box is not None else BB

    def changed(self):
        c, self._changed = self._changed, None
        return c

    def tape(self):
        raise NotImplementedError

    def oracle_sampler(self):
        return self._oracle.Sampler(params=self._oracle_params())


def flat6(bb):
    return tuple(bb[0]) + tuple(bb[1]) if len(bb) == 2 else tuple(bb)


@pytest.mark.parametrize("dims,passes", [((24, 20, 16), 2), ((11, 11, 11), 3), ((33, 7, 5), 1), ((2, 2, 2), 3)])
def test_host_sampled_full_load(S, oracle, dims, passes):
    sdf = HostDemo(S, oracle)
    o = oracle.Viewer(BB, dims, passes)
    want_it = o.update(sdf.oracle_sampler())
    with S.SDFViewer.new_voxels(dims, BB, passes) as v:
        total = len(v.loading_mgr)
        it = v.update_surface(sdf, max_delta_time=3600.0)
        assert it == want_it == total
        assert len(v.loading_mgr) == 0 and v.loading_mgr.passes_left() == 0
        assert v.loading_mgr.total_iterations() == total
        t0, t1 = v.download()
        assert_same_volume(t0, t1, o.tex0, o.tex1)
        # every voxel was AIR_DIST exactly once: sampled once, never again (scene/sdf/mod.rs:184)
        assert sdf.points == dims[0] * dims[1] * dims[2]
        assert v.update_surface(sdf, 3600.0) == 0


def test_host_sampled_time_budget_and_resume(S, oracle):
    """max_delta_time = 0 still performs one iteration (scene/sdf/mod.rs:173); a pass may stop half way
    and the next call continues at the cursor.  After every call the volume equals the oracle's after
    the same number of iterations."""
    dims = (20, 12, 10)
    sdf = HostDemo(S, oracle)
    o = oracle.Viewer(BB, dims, 2)
    s = sdf.oracle_sampler()
    with S.SDFViewer.new_voxels(dims, BB, 2) as v:
        total = len(v.loading_mgr)
        assert v.update_surface(sdf, 0.0) == 1
        assert o.update(s, max_iterations=1) == 1
        t0, t1 = v.download()
        assert_same_volume(t0, t1, o.tex0, o.tex1)
        assert len(v.loading_mgr) == o.len() == total - 1
        done, calls = 1, 0
        while len(v.loading_mgr):
            it = v.update_surface(sdf, 1e-4)   # a budget that ends inside a pass
            assert it >= 1
            assert o.update(s, max_iterations=it) == it
            done += it
            calls += 1
            assert len(v.loading_mgr) == o.len() == total - done
            assert v.loading_mgr.passes_left() == o.passes_left()
            if calls % 3 == 1:
                t0, t1 = v.download()
                assert_same_volume(t0, t1, o.tex0, o.tex1)
        assert done == total
        t0, t1 = v.download()
        assert_same_volume(t0, t1, o.tex0, o.tex1)
        v.commit()
        assert v.trace_params(S.default_camera(64, 48), 64, 48)[2] == 1.0   # lod 2^0 once loading is done


def test_host_sampled_changed_box(S, oracle):
    """changed() -> 3-pass re-sample of the voxels inside the merged box (scene/sdf/mod.rs:131-154),
    reported after and DURING a load, with the host-sampled path."""
    dims = (24, 20, 16)
    sdf = HostDemo(S, oracle)
    o = oracle.Viewer(BB, dims, 2)
    with S.SDFViewer.new_voxels(dims, BB, 2) as v:
        def step(budget, max_iterations=None):
            changed = sdf._changed
            it = v.update_surface(sdf, budget)
            want = o.update(sdf.oracle_sampler(), changed=flat6(changed) if changed is not None else None,
                            max_iterations=it if it else 1)
            assert it == want
            t0, t1 = v.download()
            assert_same_volume(t0, t1, o.tex0, o.tex1)
            assert len(v.loading_mgr) == o.len() and v.loading_mgr.passes_left() == o.passes_left()
            return it

        assert step(0.0) == 1                                   # load begins
        sdf.set(box=(-0.5, -0.25, -1.0, 0.25, 0.5, 0.1), sphere_radius=0.9)
        assert step(1e-4) >= 1                                  # change reported while loading
        while len(v.loading_mgr):
            step(3600.0)
        assert step(3600.0) > 0                                 # queued 3-pass re-sample (changed_box_while_loading)
        assert step(3600.0) > 0                                 # one more round, then the box is dropped
        assert step(3600.0) == 0
        before = sdf.points
        sdf.set(box=(0.1, 0.1, 0.1, 0.9, 0.6, 0.7), cube_half_side=0.7)
        assert step(3600.0) > 0                                 # change after loading: new 3-pass manager
        assert 0 < sdf.points - before < 3 * dims[0] * dims[1] * dims[2] // 8   # only the box was sampled
        step(3600.0)
        assert step(3600.0) == 0


def test_surface_with_tape_runs_on_the_gpu(S, oracle):
    """A surface that provides a tape never has sample() called; a parameter change re-sends the tape."""
    dims = (32, 32, 32)

    class Counting(S.SDFDemo):
        samples = 0

        def sample(self, points, distance_only=False):
            Counting.samples += 1
            raise AssertionError("sample() must not be called for a surface with a tape")

    sdf = Counting()
    o = oracle.Viewer(BB, dims, 2)
    with S.SDFViewer.new_voxels(dims, BB, 2) as v:
        l0 = v.launch_count
        it = v.update_surface(sdf, 0.030)
        assert it == o.update(oracle.Sampler(tape=sdf.tape())) == len(S.LoadingManager(dims, 2))
        assert v.launch_count - l0 == 2                         # one launch per LoadingManager pass
        t0, t1 = v.download()
        assert_same_volume(t0, t1, o.tex0, o.tex1)
        sdf.set_parameter("sphere_radius", 0.8)
        it = v.update_surface(sdf, 0.030)
        assert it == o.update(oracle.Sampler(tape=sdf.tape()), changed=flat6(BB))
        t0, t1 = v.download()
        assert_same_volume(t0, t1, o.tex0, o.tex1)
        assert Counting.samples == 0


def test_host_sampled_after_unknown_state_reads_tex0(S, oracle):
    """When the host cannot know which voxels hold AIR_DIST (an ingest in the middle of a load) the
    host-sampled path reads tex0.r of its candidates (the test of scene/sdf/mod.rs:184)."""
    dims = (16, 12, 10)
    n = dims[0] * dims[1] * dims[2]
    sdfA, sdfB = HostDemo(S, oracle), HostDemo(S, oracle, cube_half_side=0.6, sphere_radius=0.7)
    a = oracle.Viewer(BB, dims, 1); a.fill_all(sdfA.oracle_sampler())
    b = oracle.Viewer(BB, dims, 1); b.fill_all(sdfB.oracle_sampler())
    air = np.float32(oracle.lib().orc_air_dist())
    with S.SDFViewer.new_voxels(dims, BB, 2) as v:
        pos = v.voxel_positions(100, 500)
        v.ingest_samples(100, sdfB.sample(pos))                 # knowledge lost before the load starts
        assert not np.any(b.tex0.reshape(-1, 4)[100:600, 0] == air)
        assert v.update_surface(sdfA, 3600.0) == len(S.LoadingManager(dims, 2))
        assert sdfA.points == n - 500                           # the ingested voxels were left alone
        t0, t1 = v.download()
        want0, want1 = a.tex0.copy().reshape(-1, 4), a.tex1.copy().reshape(-1, 4)
        want0[100:600], want1[100:600] = b.tex0.reshape(-1, 4)[100:600], b.tex1.reshape(-1, 4)[100:600]
        assert_same_volume(t0, t1, want0.reshape(t0.shape), want1.reshape(t1.shape))


def test_host_sampled_slab_handle(S, oracle):
    """A slab handle samples only the slices it stores but counts every iteration of the global grid."""
    dims = (12, 10, 16)
    sdf = HostDemo(S, oracle)
    o = oracle.Viewer(BB, dims, 2); o.update(sdf.oracle_sampler())
    with S.SDFViewer.new_voxels(dims, BB, 2, z_range=(4, 9)) as v:
        assert v.update_surface(sdf, 3600.0) == len(S.LoadingManager(dims, 2))
        assert sdf.points == dims[0] * dims[1] * (v.z_hi - v.z_lo)
        t0, t1 = v.download()
        assert_same_volume(t0, t1, o.tex0[4:9], o.tex1[4:9])


def test_update_surface_errors(S, oracle):
    import ctypes as C
    from sdf_viewer_b200 import _lib
    with S.SDFViewer.new_voxels((4, 4, 4), BB, 1) as v:
        it = C.c_uint64()
        assert v._lib.sdfgpu_update_surface(v._h, None, 0.03, C.byref(it)) == _lib.SDFGPU_ERR_INVALID
        empty = _lib.Surface()
        assert v._lib.sdfgpu_update_surface(v._h, C.byref(empty), 0.03, C.byref(it)) == _lib.SDFGPU_ERR_INVALID
        assert b"neither a tape nor a sample" in v._lib.sdfgpu_last_error(v._h)

        class Broken(HostDemo):
            def sample(self, points, distance_only=False):
                raise RuntimeError("guest trapped")
        with pytest.raises(RuntimeError, match="guest trapped"):
            v.update_surface(Broken(S, oracle), 3600.0)
        # the library itself returned the reference's benign sample (distance 1.0, native.rs:202) and went on
        t0, _ = v.download()
        assert np.all(t0[..., 0] == 1.0)
