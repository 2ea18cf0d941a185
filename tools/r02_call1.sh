#!/bin/bash
# Round 2, first gpurun call (one B200): the tests that were still marked gpu_next, the configurations that had
# no kept evidence (C3 CSG-1k, C5 dirty re-sample, the WebAssembly guest), and ncu captures of those fill kernels.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader > gpurun_out/r02_gpu.txt 2>&1
SDFGPU_RUN_NEXT=1 timeout 400 python -m pytest tests -q -m gpu_next > gpurun_out/r02_gpu_next_tests.log 2>&1
echo "gpu_next tests rc=$?"; tail -n 3 gpurun_out/r02_gpu_next_tests.log
timeout 240 python tools/configs_run.py csg dirty trace > gpurun_out/r02_configs.txt 2>&1
echo "configs rc=$?"; cat gpurun_out/r02_configs.txt
for v in 1 2 4 8; do
  timeout 120 python bench.py --workload wasm --vpt $v --steps 30 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_wasm_v$v.json 2> gpurun_out/r02_bench_wasm_v$v.err
  echo "bench wasm vpt $v rc=$?"; python -c "import json;d=json.load(open('gpurun_out/r02_bench_wasm_v$v.json'));print(' fill_ms',d['fill_ms'],'samples/s',d['fill_samples_per_sec'])"
done
timeout 120 python bench.py --workload csg --steps 30 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_csg.json 2> gpurun_out/r02_bench_csg.err
echo "bench csg rc=$?"; tail -c 400 gpurun_out/r02_bench_csg.json
timeout 200 ncu --set full --clock-control none --import-source on -k regex:sdfgpu_fill_jit -s 1 -c 1 -f -o gpurun_out/r02_fill_wasm_demo \
  python bench.py --workload wasm --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r02_ncu_fill_wasm.log 2>&1
echo "ncu wasm fill rc=$?"
timeout 200 ncu --set full --clock-control none --import-source on -k regex:sdfgpu_fill_jit -s 1 -c 1 -f -o gpurun_out/r02_fill_csg \
  python tools/profile_run.py 512 csg 2 > gpurun_out/r02_ncu_fill_csg.log 2>&1
echo "ncu csg fill rc=$?"
exit 0
