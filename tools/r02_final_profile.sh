#!/bin/bash
# round 2, final build, one GPU: the GPU suite, smoke, the bench's launch list, full ncu sets of the hot kernels
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -x -q -m gpu > gpurun_out/r02z_gpu_tests.log 2>&1; echo "gpu tests rc=$?"; tail -3 gpurun_out/r02z_gpu_tests.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02z_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r02z_smoke.log
NCU="ncu --clock-control none"
$NCU --metrics gpu__time_duration.sum -c 400 --csv --log-file gpurun_out/r02z_launches_bench_steps3.csv python bench.py --steps 3 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/r02z_launches_bench.log 2>&1; echo "launch list rc=$?"
$NCU --set full --import-source on -k regex:fill -s 1 -c 1 -o gpurun_out/r02z_fill -f python tools/profile_run.py 512 demo 3 > gpurun_out/r02z_ncu_fill.log 2>&1; echo "fill rc=$?"
$NCU --set full --import-source on -k regex:trace_tiles -s 1 -c 1 -o gpurun_out/r02z_trace -f python tools/profile_run.py 512 demo 3 > gpurun_out/r02z_ncu_trace.log 2>&1; echo "trace rc=$?"
$NCU --set full --import-source on -k regex:"fill|cull_cells" -c 3 -o gpurun_out/r02z_fill_csg -f python tools/profile_run.py 512 csg 2 > gpurun_out/r02z_ncu_fill_csg.log 2>&1; echo "csg rc=$?"
for n in r02z_fill r02z_trace r02z_fill_csg; do
  ncu -i gpurun_out/$n.ncu-rep --page raw --csv > gpurun_out/${n}_ncu_raw.csv 2>/dev/null
  python tools/ncu_summary.py gpurun_out/${n}_ncu_raw.csv > gpurun_out/${n}_ncu_summary.txt 2>&1
  grep "Kernel Name\|gpu__time_duration\|dram__bytes\|registers_per_thread\|issue_active\|sm__throughput\|l1tex__t_sector_hit\|lts__t_sector_hit\|cycles_elapsed.max\|cycles_active.avg\|inst_executed.sum" gpurun_out/${n}_ncu_summary.txt
done
head -12 gpurun_out/r02z_launches_bench_steps3.csv | cut -c1-220
