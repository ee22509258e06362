#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/time_conv.py 8,256,256,256,256,9 8,256,256,256,256,9,0,2 8,128,128,512,512,9,0,2 > gpurun_out/time_conv.log 2>&1
cat gpurun_out/time_conv.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_tc_persist -c 4 -o gpurun_out/r01_conv_mode2 python tools/time_conv.py 8,256,256,256,256,9,0,2 > gpurun_out/ncu_mode2.log 2>&1
