#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_linked_gpu.py tests/test_trace_gpu.py -x -q 2>&1 | tail -8
python tools/link_timing.py 2>&1 | tail -14
python tools/trace_tail.py 2>&1 | grep "maxSteps 256\|histogram"
