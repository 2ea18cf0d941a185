#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_linked_gpu.py tests/test_trace_gpu.py -x -q 2>&1 | tail -12
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29541 tests/multi_gpu_check.py 2>&1 | grep -v "^\*\*\*\|OMP_NUM" | tail -12
for A in 1 0; do
LINK_ASSIGN=$A SDFGPU_LINK_TIMING=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 2956$A tools/link_timing.py > gpurun_out/r02h_timing_n2_a$A.log 2>&1; echo "assign $A rc=$?"; grep -v "^\*\*\*\|OMP_NUM" gpurun_out/r02h_timing_n2_a$A.log | grep "==\|frame 2 \|frame 25\|frame 48\|frame 71"
done
