#!/bin/bash
# One gpurun call: the new host-sampled / C++ host tests first, then a short bench, then the whole GPU suite.
# Everything is logged under gpurun_out/ so a cut-off call still leaves evidence.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader > gpurun_out/gpu.txt 2>&1
timeout 150 python -m pytest tests/test_update_surface_gpu.py tests/test_cpp_host.py -x -q -m gpu > gpurun_out/new_tests.log 2>&1
echo "new tests rc=$?" | tee -a gpurun_out/new_tests.log
timeout 120 python bench.py --steps 50 --warmup 5 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
echo "bench rc=$?"
tail -c 600 gpurun_out/bench_n1.json
timeout 400 python -m pytest tests -x -q -m gpu > gpurun_out/gpu_tests.log 2>&1
echo "gpu suite rc=$?" | tee -a gpurun_out/gpu_tests.log
tail -5 gpurun_out/new_tests.log gpurun_out/gpu_tests.log
