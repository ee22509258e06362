#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/time_conv.py 1,256,256,256,256,9 1,256,256,256,256,9,1 1,256,256,256,256,9,0,2 8,256,256,256,256,9 8,256,256,256,256,9,1 8,256,256,256,256,9,0,1 8,256,256,256,256,9,0,2 8,128,128,512,512,9 8,128,128,512,512,9,0,2 > gpurun_out/time_conv.log 2>&1
cat gpurun_out/time_conv.log
for f in 1 2; do
  OSM_GN_FUSE=$f timeout 300 python tools/profile_step.py --batch 1 > gpurun_out/step_b1_fuse$f.log 2> gpurun_out/step_b1.err
  echo "FUSE=$f B=1"; head -2 gpurun_out/step_b1_fuse$f.log
  OSM_GN_FUSE=$f timeout 300 python tools/profile_step.py --batch 8 > gpurun_out/step_b8_fuse$f.log 2>&1
  echo "FUSE=$f B=8"; grep "^step" gpurun_out/step_b8_fuse$f.log
done
timeout 600 python -m pytest tests -m gpu -x -q -k "conv or unet or step" > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
