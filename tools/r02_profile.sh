#!/bin/bash
# round 2 ncu captures (one GPU): launch list of the bench, full sets of the new kernels
mkdir -p gpurun_out
NCU="ncu --clock-control none"
$NCU --metrics gpu__time_duration.sum -c 400 --csv --log-file gpurun_out/r02_launches_bench_steps3.csv python bench.py --steps 3 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/r02_launches_bench.log 2>&1; echo "launch list rc=$?"
$NCU --set full --import-source on -k regex:mesh_ -c 5 -o gpurun_out/r02_mesh -f python tools/profile_run2.py mesh 512 > gpurun_out/r02_ncu_mesh.log 2>&1; echo "mesh rc=$?"
$NCU --set full --import-source on -k regex:trace_rounds -s 2 -c 2 -o gpurun_out/r02_trace_rounds -f python tools/profile_run2.py linked 512 > gpurun_out/r02_ncu_linked.log 2>&1; echo "linked rc=$?"
$NCU --set full --import-source on -k regex:fill -s 2 -c 1 -o gpurun_out/r02_fill_linked -f python tools/profile_run2.py linked 512 > gpurun_out/r02_ncu_fill_linked.log 2>&1; echo "fill linked rc=$?"
$NCU --set full --import-source on -k regex:fill -s 1 -c 1 -o gpurun_out/r02_fill_points -f python tools/profile_run2.py points > gpurun_out/r02_ncu_points.log 2>&1; echo "points rc=$?"
for n in r02_mesh r02_trace_rounds r02_fill_linked r02_fill_points; do
  ncu -i gpurun_out/$n.ncu-rep --page raw --csv > gpurun_out/${n}_raw.csv 2>/dev/null
  python tools/ncu_summary.py gpurun_out/${n}_raw.csv > gpurun_out/${n}_ncu_summary.txt 2>&1
  grep "Kernel Name\|gpu__time_duration\|dram__bytes\|registers_per_thread\|issue_active\|l1tex__t_sector_hit\|lts__t_sector_hit\|cycles_elapsed.max\|cycles_active.avg" gpurun_out/${n}_ncu_summary.txt
done
head -30 gpurun_out/r02_launches_bench_steps3.csv | cut -c1-200
