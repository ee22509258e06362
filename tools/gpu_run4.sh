#!/bin/bash
# GPU pass 4: straight-line conv epilogue; fused statistics cost
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 300 python tools/time_conv.py 1,256,256,256,256,9 1,256,256,256,256,9,1 1,256,256,256,256,9,0,1 1,256,256,256,256,9,0,2 8,256,256,256,256,9 8,256,256,256,256,9,1 8,256,256,256,256,9,0,1 8,256,256,256,256,9,0,2 8,128,128,512,512,9 8,128,128,512,512,9,0,1 8,128,128,512,512,9,0,2 1,128,128,256,256,9 1,128,128,256,256,9,0,1 > gpurun_out/time_conv.log 2>&1
cat gpurun_out/time_conv.log
timeout 300 python tools/profile_step.py --batch 1 > gpurun_out/step_b1.log 2> gpurun_out/step_b1.err
head -5 gpurun_out/step_b1.log
OSM_GN_FUSE=0 timeout 300 python tools/profile_step.py --batch 1 > gpurun_out/step_b1_nofuse.log 2> gpurun_out/step_b1.err
head -5 gpurun_out/step_b1_nofuse.log
timeout 300 python tools/profile_step.py --batch 8 > gpurun_out/step_b8.log 2>&1
head -5 gpurun_out/step_b8.log
OSM_GN_FUSE=0 timeout 300 python tools/profile_step.py --batch 8 > gpurun_out/step_b8_nofuse.log 2>&1
head -5 gpurun_out/step_b8_nofuse.log
