#!/bin/bash
# First gpurun call of the next round (one B200): everything DESIGN.md section 8 lists for a single GPU.
#   gpurun --timeout 900 -- 'bash tools/gpu_next_round.sh'
mkdir -p gpurun_out
SDFGPU_RUN_NEXT=1 timeout 300 python -m pytest tests -q -m gpu_next > gpurun_out/gpu_next_tests.log 2>&1
echo "gpu_next tests rc=$?"; tail -n 3 gpurun_out/gpu_next_tests.log
timeout 300 python -m pytest tests -x -q -m gpu > gpurun_out/gpu_tests.log 2>&1
echo "gpu suite rc=$?"; tail -n 2 gpurun_out/gpu_tests.log
timeout 200 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
echo "bench rc=$?"
for v in 1 2 4 8; do
  timeout 120 python bench.py --workload wasm --vpt $v --steps 30 --warmup 3 --no-cpu-baseline > gpurun_out/bench_wasm_v$v.json 2> gpurun_out/bench_wasm_v$v.err
  echo "bench wasm vpt $v rc=$?"; python -c "import json;d=json.load(open('gpurun_out/bench_wasm_v$v.json'));print(' fill_ms',d['fill_ms'],'samples/s',d['fill_samples_per_sec'])"
done
timeout 200 ncu --set full --clock-control none --import-source on -k regex:sdfgpu_fill_jit -s 1 -c 1 -f -o gpurun_out/fill_wasm_demo \
  python bench.py --workload wasm --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_fill_wasm.log 2>&1
echo "ncu wasm fill rc=$?"
timeout 120 python tools/trace_modes.py 512 > gpurun_out/trace_modes.txt 2>&1; cat gpurun_out/trace_modes.txt
exit 0
