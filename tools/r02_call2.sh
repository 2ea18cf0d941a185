#!/bin/bash
# round 2, GPU call 2: linked slabs -- single-device group tests, then real peers (2 ranks)
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r02b_gpu.txt
timeout 600 python -m pytest tests/test_linked_gpu.py -x -q > gpurun_out/r02b_linked_tests.log 2>&1
echo "linked tests rc=$?"; tail -15 gpurun_out/r02b_linked_tests.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29541 tests/multi_gpu_check.py > gpurun_out/r02b_multi_gpu_check.log 2>&1
echo "multi_gpu_check rc=$?"; tail -25 gpurun_out/r02b_multi_gpu_check.log
