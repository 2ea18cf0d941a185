#!/bin/bash
# round 2, N = 2: linked tests (both halo modes, stream + rounds trace), link timing, a short bench
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_linked_gpu.py tests/test_sharded_gpu.py -x -q -m gpu > gpurun_out/r02i_linked_tests.log 2>&1
echo "tests rc=$?"; tail -5 gpurun_out/r02i_linked_tests.log
for mode in "0 0" "0 1" "1 0"; do
  set -- $mode
  SDFGPU_LINK_TIMING=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29581 tools/link_timing.py 512 $1 $2 > gpurun_out/r02i_timing_n2_h$1_t$2.log 2>&1
  echo "timing halo_push=$1 trace_mode=$2 rc=$?"; grep "^==" gpurun_out/r02i_timing_n2_h$1_t$2.log
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29582 bench.py --gpus 2 --steps 30 --warmup 5 > gpurun_out/r02i_bench_n2.json 2> gpurun_out/r02i_bench_n2.err
echo "bench rc=$?"; tail -c 3000 gpurun_out/r02i_bench_n2.json; tail -5 gpurun_out/r02i_bench_n2.err
