#!/bin/bash
# GPU pass 2: m256 conv kernel, model-based conv policy, GN tail fix, new loop test
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
OSM_CONV_M256=2 timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "conv" > gpurun_out/pytest_m256.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_m256.log
tail -5 gpurun_out/pytest_m256.log
OSM_CONV_VERBOSE=1 timeout 300 python tools/time_conv.py > gpurun_out/time_conv.log 2>&1
OSM_CONV_M256=0 timeout 300 python tools/time_conv.py 1,256,256,256,256,9 8,256,256,256,256,9 8,128,128,512,512,9 1,256,256,256,512,9 > gpurun_out/time_conv_nom256.log 2>&1
cat gpurun_out/time_conv.log | grep -v conv_tc_plan
cat gpurun_out/time_conv_nom256.log
OSM_CONV_VERBOSE=1 timeout 300 python tools/profile_step.py --batch 1 > gpurun_out/step_b1.log 2> gpurun_out/step_b1.err
head -5 gpurun_out/step_b1.log
timeout 300 python tools/profile_step.py --batch 8 > gpurun_out/step_b8.log 2>&1
head -5 gpurun_out/step_b8.log
timeout 600 python bench.py --steps 30 --warmup 3 --no-cpu-baseline > gpurun_out/bench_b1.json 2> gpurun_out/bench_b1.err
tail -2 gpurun_out/bench_b1.json | cut -c1-400
