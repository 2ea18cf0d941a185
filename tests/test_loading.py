"""The reference's own unit tests for the fill order, ported verbatim
(/root/reference/src/app/scene/sdf/loading.rs:117-171, five cases, 3 passes): every voxel is
yielded at least once and at most `passes` times, and iterations_done + len() is invariant.
Run against BOTH the oracle's LoadingManager (C++) and the host mirror
(sdf-viewer_b200/loading.py) -- these are the only golden facts the reference pins for this path."""
import numpy as np
import pytest

CASES = [(2, 2, 2), (8, 8, 8), (64, 64, 64), (11, 11, 11), (8, 11, 17)]  # loading.rs:147-170


def _check(limits, next_fn, len_fn, num_passes=3):
    hits = np.zeros(limits[0] * limits[1] * limits[2], np.int64)
    remaining = len_fn()
    iterations = 0
    total = iterations + remaining
    while True:
        v = next_fn()
        if v is None:
            break
        flat = v[0] + v[1] * limits[0] + v[2] * limits[0] * limits[1]
        hits[flat] += 1
        assert hits[flat] <= num_passes
        iterations += 1
        assert total == iterations + len_fn()
    assert hits.min() >= 1, f"voxel not hit: {np.argmin(hits)}"
    return iterations


@pytest.mark.parametrize("limits", CASES)
def test_reference_cases_oracle(oracle, limits):
    lm = oracle.LM(limits, 3)
    n = _check(limits, lm.next, lm.len)
    assert n == lm.total_iterations() and lm.passes_left() == 0


@pytest.mark.parametrize("limits", CASES)
def test_reference_cases_host_mirror(S, limits):
    lm = S.LoadingManager(limits, 3)

    def nxt():
        try:
            return next(lm)
        except StopIteration:
            return None
    n = _check(limits, nxt, lambda: len(lm))
    assert n == lm.total_iterations() and lm.passes_left() == 0


@pytest.mark.parametrize("limits,passes", [((5, 3, 7), 0), ((5, 3, 7), 1), ((9, 4, 6), 2), ((16, 16, 3), 4), ((3, 3, 3), 6)])
def test_host_mirror_equals_oracle_sequence(S, oracle, limits, passes):
    """Same visit sequence, len(), passes_left() at every step (loading.rs:50-105)."""
    a, b = oracle.LM(limits, passes), S.LoadingManager(limits, passes)
    assert a.len() == len(b) and a.passes_left() == b.passes_left()
    while True:
        va = a.next()
        try:
            vb = next(b)
        except StopIteration:
            vb = None
        assert va == vb
        assert a.len() == len(b) and a.passes_left() == b.passes_left()
        assert a.total_iterations() == b.total_iterations()
        if va is None:
            break


def test_pass_structure(S):
    """A pass with step s visits exactly the lattice {0, s, 2s, ...}^3, x fastest (the launch
    geometry of the fill kernel) and SURVEY 8a row 4's iteration counts."""
    L = S.loading
    assert L.pass_steps(2) == [2, 1] and L.pass_steps(3) == [4, 2, 1] and L.pass_steps(0) == [1] and L.pass_steps(1) == [1]
    assert sum(L.pass_items((64,) * 3, s) for s in L.pass_steps(2)) == 294912
    assert sum(L.pass_items((512,) * 3, s) for s in L.pass_steps(3)) == 153_092_096
    lm = L.LoadingManager((5, 4, 3), 2)
    first = [next(lm) for _ in range(L.pass_items((5, 4, 3), 2))]
    want = [(x, y, z) for z in range(0, 3, 2) for y in range(0, 4, 2) for x in range(0, 5, 2)]
    assert first == want and lm.step_size == 1
    for x in (0, 1, 2, 3, 5, 8, 9, 1023, 1024, 2 ** 31 + 5):
        assert L.prev_power_of_2(x) == (0 if x == 0 else 1 << (x.bit_length() - 1))


# ---- the library's own LoadingManager object (sdfgpu_loading_*, device-free): the state every handle keeps

@pytest.mark.parametrize("limits", CASES)
def test_reference_cases_native(S, limits):
    lm = S.NativeLoadingManager(limits, 3)

    def nxt():
        try:
            return next(lm)
        except StopIteration:
            return None
    n = _check(limits, nxt, lambda: len(lm))
    assert n == lm.total_iterations() and lm.passes_left() == 0


@pytest.mark.parametrize("limits,passes", [((5, 3, 7), 0), ((5, 3, 7), 1), ((9, 4, 6), 2), ((16, 16, 3), 4), ((3, 3, 3), 6),
                                           ((1, 1, 1), 3), ((33, 7, 5), 3)])
def test_native_equals_oracle_sequence(S, oracle, limits, passes):
    a, b = oracle.LM(limits, passes), S.NativeLoadingManager(limits, passes)
    assert a.len() == len(b) and a.passes_left() == b.passes_left()
    while True:
        va = a.next()
        try:
            vb = next(b)
        except StopIteration:
            vb = None
        assert va == vb
        assert a.len() == len(b) and a.passes_left() == b.passes_left()
        assert a.total_iterations() == b.total_iterations()
        if va is None:
            break


@pytest.mark.parametrize("limits,passes,chunk", [((5, 3, 7), 2, 1), ((9, 4, 6), 3, 2), ((16, 16, 3), 4, 5), ((33, 7, 5), 3, 1000),
                                                 ((64, 64, 64), 2, 37)])
def test_native_runs_concatenate_to_the_reference_order(S, limits, passes, chunk):
    """next_run -- what the host-sampled update walks in chunks (sdfgpu_update_surface) -- yields the
    same sequence, len() and counters as one next() at a time."""
    one, run = S.LoadingManager(limits, passes), S.NativeLoadingManager(limits, passes)
    while True:
        r = run.next_run(chunk)
        if r is None:
            assert len(one) == 0 and run.passes_left() == 0
            with pytest.raises(StopIteration):
                next(one)
            break
        (x0, y, z), step, n = r
        assert 1 <= n <= chunk
        for i in range(n):
            assert next(one) == (x0 + i * step, y, z)
        assert len(one) == len(run) and one.total_iterations() == run.total_iterations()
        assert one.passes_left() == run.passes_left()


def test_native_reset(S):
    lm = S.NativeLoadingManager((4, 4, 4), 2)
    for _ in range(5):
        next(lm)
    lm.reset(3)
    ref = S.LoadingManager((4, 4, 4), 3)
    assert len(lm) == len(ref) and lm.total_iterations() == 0 and lm.passes_left() == 3
    assert [next(lm) for _ in range(len(ref))] == [next(ref) for _ in range(len(ref))]
