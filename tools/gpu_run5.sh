#!/bin/bash
# GPU pass 5: post-processing kernels, record, reverse-order GN apply
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -25 gpurun_out/pytest_gpu.log
timeout 300 python tools/profile_step.py --batch 1 > gpurun_out/step_b1.log 2> gpurun_out/step_b1.err
head -3 gpurun_out/step_b1.log
timeout 300 python tools/profile_step.py --batch 8 > gpurun_out/step_b8.log 2>&1
head -5 gpurun_out/step_b8.log | tail -3
