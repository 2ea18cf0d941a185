#!/bin/bash
# Round-end rehearsal on one B200: whole GPU suite, smoke(), bench, launch list of a short bench run.
mkdir -p gpurun_out
timeout 200 python -m pytest tests -x -q -m gpu > gpurun_out/gpu_tests_final.log 2>&1
echo "gpu suite rc=$?" | tee -a gpurun_out/gpu_tests_final.log
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
timeout 120 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err
echo "bench rc=$?"
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench_steps3.csv python bench.py --steps 3 --warmup 1 --no-cpu-baseline > gpurun_out/launches.log 2>&1
echo "launch list rc=$?"
tail -n 3 gpurun_out/gpu_tests_final.log
exit 0
