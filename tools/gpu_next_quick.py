"""The shortest possible first GPU run of a scalar-program tape and of a lowered WebAssembly guest (bit-exact
against the oracle).  The full versions are the gpu_next tests (SDFGPU_RUN_NEXT=1 pytest -m gpu_next)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import orc
import sdf_viewer_b200 as S
import test_scalar_programs as P
import test_wasm_lower as W

BB = ((-1.0, -1.0, -1.0), (1.0, 1.0, 1.0))
dims = (40, 36, 32)
P.NAMES.update({code: name for name, code in S.tape.S.items()})
tapes = {"scalar sphere_with_bands": P.build_tape(S.tape, P.sphere_with_bands(S.tape))[1],
         "wasm csg_calls": S.wasm.lower(W.guest_csg_calls().build())[0],
         "wasm early_returns": S.wasm.lower(W.guest_early_returns().build())[0]}
for name, tape in tapes.items():
    o = orc.Viewer(BB, dims, 2)
    o.update(orc.Sampler(tape=tape))
    with S.SDFViewer.new_voxels(dims, BB, 2) as v:
        v.set_tape(tape)
        v.update(None)
        t0, t1 = v.download()
        prog = v.get_info("last_fill_program")
    ok = P.same_f32(t0, o.tex0) and P.same_f32(t1, o.tex1)
    print(f"{name}: program {prog}, bit-exact {ok}", flush=True)
    assert ok and prog == 1
print("gpu_next_quick ok")
