#!/bin/bash
# round 2, N = 1: fill tests (coarse-cell pre-cull), CSG timing with and without the pre-cull
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_fill_gpu.py tests/test_full_size_gpu.py -x -q -m gpu > gpurun_out/r02m_fill_tests.log 2>&1
echo "tests rc=$?"; tail -3 gpurun_out/r02m_fill_tests.log
python - <<'PY'
import sys, time, torch
sys.path.insert(0, ".")
import sdf_viewer_b200 as S
BB = ((-1.0, -1.0, -1.0), (1.0, 1.0, 1.0))
tape = S.tape.csg_tape()
for cells in (1, 0):
    for vpt in (8, 4):
        with S.SDFViewer.new_voxels((512, 512, 512), BB, 1) as v:
            v.set_option("fill_cull_cells", cells); v.set_option("fill_voxels_per_thread", vpt)
            t = time.perf_counter(); v.set_tape(tape); v.sync(); ts = time.perf_counter() - t
            t = time.perf_counter(); v.set_tape(tape); v.sync(); ts2 = time.perf_counter() - t
            stream = torch.cuda.ExternalStream(v.stream)
            for _ in range(3): v.fill_all()
            v.sync()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for _ in range(10): v.fill_all()
            e1.record(stream); v.sync(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 10
            print(f"csg-1k 512^3 cull_cells={cells} vpt={vpt}: {ms:.3f} ms {512**3/ms/1e6:.1f} Gsamples/s; set_tape first {ts*1e3:.2f} ms, again {ts2*1e3:.2f} ms; {v.cull_stats()}", flush=True)
PY
