#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_mesh.py tests/test_linked_gpu.py tests/test_trace_gpu.py tests/test_fill_gpu.py -x -q > gpurun_out/r02f_tests.log 2>&1; echo "tests rc=$?"; tail -25 gpurun_out/r02f_tests.log
python tools/link_timing.py > gpurun_out/r02e_timing_n1.log 2>&1; echo "n1 rc=$?"; tail -14 gpurun_out/r02e_timing_n1.log
python - <<'PY'
import sys, time
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np, torch
import sdf_viewer_b200 as S
BB = ((-1.0, -1.0, -1.0), (1.0, 1.0, 1.0))
for side in (256, 512):
    with S.SDFViewer.from_bb(BB, side, 1) as v:
        v.set_tape(S.tape.demo_tape()); v.fill_all(); v.sync()
        v.mesh(download=False); v.sync()
        t = time.perf_counter()
        for _ in range(5):
            nv, nt = v.mesh(download=False)
        v.sync()
        dt = (time.perf_counter() - t) / 5
        print(f"mesh {side}^3: {nv} vertices {nt} triangles, {dt*1e3:.3f} ms -> {nt/dt/1e6:.1f} Mtri/s, {(side-1)**3/dt/1e9:.2f} Gcells/s", flush=True)
PY
