#!/bin/bash
# round 2, N = 2: link timing of the stream trace (default halo mode) + multi-GPU checks
mkdir -p gpurun_out
SDFGPU_LINK_TIMING=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29581 tools/link_timing.py 512 0 0 > gpurun_out/r02j_timing_n2.log 2>&1
echo "timing rc=$?"; grep "^==" gpurun_out/r02j_timing_n2.log
timeout 900 python -m pytest tests/test_sharded_gpu.py -x -q -m gpu > gpurun_out/r02j_sharded_tests.log 2>&1
echo "tests rc=$?"; tail -3 gpurun_out/r02j_sharded_tests.log
