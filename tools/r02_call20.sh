#!/bin/bash
# round 2: GPU suite + smoke + the N = 1 bench line of the final build
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -x -q -m gpu > gpurun_out/r02y_gpu_tests.log 2>&1; echo "gpu tests rc=$?"; tail -3 gpurun_out/r02y_gpu_tests.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02y_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r02y_smoke.log
bash tools/r02_scale_one.sh 1
