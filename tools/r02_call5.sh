#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_linked_gpu.py -x -q 2>&1 | tail -3
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29541 tests/multi_gpu_check.py 2>&1 | grep -v "^\*\*\*\|OMP_NUM" | tail -4
SDFGPU_LINK_TIMING=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29561 tools/link_timing.py > gpurun_out/r02e_timing_n2.log 2>&1; echo "n2 rc=$?"; grep -v "^\*\*\*\|OMP_NUM" gpurun_out/r02e_timing_n2.log | grep "==\|frame 2 \|frame 25\|frame 48\|frame 71"
