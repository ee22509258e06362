#!/bin/bash
# round-1 session-3 GPU pass 1: flash attention bring-up + small-conv diagnosis
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv > gpurun_out/gpu.txt
timeout 300 python tools/test_flash.py > gpurun_out/flash_test.log 2>&1; echo "flash_test exit $?" >> gpurun_out/flash_test.log
tail -12 gpurun_out/flash_test.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -8 gpurun_out/pytest_gpu.log
timeout 300 python tools/profile_step.py --batch 1 > gpurun_out/step_b1.log 2>&1
head -4 gpurun_out/step_b1.log
for f in "" "128,8" "64,8" "32,8" "64,4" "128,4" "32,4" "256,8" ; do
  echo "== OSM_CONV_FORCE=$f" >> gpurun_out/conv_small.log
  OSM_CONV_FORCE=$f timeout 120 python tools/time_conv.py 1,8,8,1024,1024,9 1,16,16,1024,1024,9 1,32,32,512,512,9 1,8,8,1024,3072,1 1,16,16,1024,1024,1 1,32,32,512,1536,1 1,64,64,512,512,9 >> gpurun_out/conv_small.log 2>&1
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_tc -o gpurun_out/r01_conv_small python tools/ncu_conv.py 1,8,8,1024,1024,9 1,16,16,1024,1024,9 1,32,32,512,512,9 > gpurun_out/ncu_conv_small.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"gn_|flash_" -o gpurun_out/r01_gn_flash python tools/ncu_misc.py > gpurun_out/ncu_misc.log 2>&1
timeout 600 python bench.py --steps 30 --warmup 3 --no-cpu-baseline > gpurun_out/bench_b1.json 2> gpurun_out/bench_b1.err
tail -2 gpurun_out/bench_b1.json
