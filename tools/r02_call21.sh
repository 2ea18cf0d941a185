#!/bin/bash
# round 2: the multi-GPU checks on 4 devices (what -m gpu runs on a box with >= 4 GPUs)
mkdir -p gpurun_out
timeout 500 python -m pytest tests/test_sharded_gpu.py -x -q -m gpu > gpurun_out/r02t_sharded_tests_n4.log 2>&1
echo "tests rc=$?"; tail -15 gpurun_out/r02t_sharded_tests_n4.log
