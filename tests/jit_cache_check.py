"""Run as a script with SDFGPU_JIT_CACHE=2 (tests/test_fill_gpu.py::test_jit_cache_is_bounded launches it): more tape
STRUCTURES than the cache of specialised kernels may hold are used in turn, twice over; evicted kernels are unloaded
(after a device synchronisation), recompiled when they come back, and every fill stays bit-identical to the oracle."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import orc  # noqa: E402
import sdf_viewer_b200 as S  # noqa: E402

BB = ((-1.0, -1.0, -1.0), (1.0, 1.0, 1.0))


def structures(T):
    out = [T.demo_tape()]
    for n in (1, 2, 3):  # n spheres united one by one: n different opcode sequences
        t = T.TapeBuilder()
        for k in range(n):
            i = t.prim(T.SHAPE_SPHERE, (0.3 * k - 0.3, 0.1 * k, 0.0), 0.4, T.MAT_NORMAL)
            t.emit(T.OP_PRIM if k == 0 else T.OP_UNION_PRIM, i)
        t.emit(T.OP_END)
        out.append(t.build())
    return out


def main():
    assert os.environ.get("SDFGPU_JIT_CACHE") == "2"
    orc.build()
    dims = (40, 36, 24)
    tapes = structures(S.tape)
    want = []
    for tape in tapes:
        o = orc.Viewer(BB, dims, 1)
        o.fill_all(orc.Sampler(tape=tape))
        want.append((o.tex0.copy(), o.tex1.copy()))
    with S.SDFViewer.new_voxels(dims, BB, 1) as v:
        v.set_option("fill_program", 3)  # the specialised kernel or fail
        for rnd in range(2):
            for k, tape in enumerate(tapes):
                v.set_tape(tape)
                v.fill_all()
                t0, t1 = v.download()
                assert np.array_equal(t0.view(np.uint32), want[k][0].view(np.uint32)), (rnd, k)
                assert np.array_equal(t1.view(np.uint32), want[k][1].view(np.uint32)), (rnd, k)
    print("jit_cache_check ok")


if __name__ == "__main__":
    main()
