#!/bin/bash
# Second gpurun call of the session: changed tests, the trace distance-source sweep, the host-sampled path
# rate per thread count, and one ncu capture of the TMU (point mode) trace.
mkdir -p gpurun_out
timeout 120 python -m pytest tests/test_update_surface_gpu.py tests/test_cpp_host.py tests/test_trace_gpu.py -x -q -m gpu > gpurun_out/new_tests2.log 2>&1
echo "changed tests rc=$?" | tee -a gpurun_out/new_tests2.log
timeout 90 python tools/trace_modes.py 512 > gpurun_out/trace_modes.txt 2>&1
echo "trace_modes rc=$?"; cat gpurun_out/trace_modes.txt
timeout 60 python - > gpurun_out/host_path.txt 2>&1 <<'PY'
import sys; sys.path.insert(0, "tests")
import bench, orc, sdf_viewer_b200 as S
orc.build()
for t in (1, 4, 16, orc.lib().orc_max_threads()):
    print(bench.host_sampled_path(orc, S, t), flush=True)
PY
echo "host_path rc=$?"; cat gpurun_out/host_path.txt
SDFGPU_TRACE_DIST=2 timeout 120 ncu --set full --clock-control none --import-source on -k regex:trace -s 1 -c 1 -f -o gpurun_out/r01_trace_tmu_point python tools/profile_run.py 512 demo 2 > gpurun_out/ncu_tmu.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/ncu_tmu.log
tail -n 5 gpurun_out/new_tests2.log
exit 0
