#!/bin/bash
# 16-CTA-cluster split-K experiment + GroupNorm reduction grid: kernel timings, parity, step time with and without
mkdir -p gpurun_out
SH="1,8,8,1024,1024,9 1,16,16,1024,1024,9 1,32,32,512,512,9 1,8,8,1024,3072,1 1,16,16,1024,1024,1 1,8,8,2048,1024,9 1,16,16,2048,1024,9 1,32,32,1024,512,9 1,64,64,512,512,9"
echo "== policy max split 8"; OSM_CONV_MAX_SPLIT=8 OSM_CONV_VERBOSE=1 timeout 300 python tools/time_conv.py $SH 2>&1 | grep -E "us|split" | sed 's/conv_tc_plan: //' 
echo "== policy max split 16"; OSM_CONV_MAX_SPLIT=16 OSM_CONV_VERBOSE=1 timeout 300 python tools/time_conv.py $SH 2>&1 | grep -E "us|split" | sed 's/conv_tc_plan: //'
for f in 128,16 256,16 64,16; do echo "== forced $f"; OSM_CONV_FORCE=$f timeout 300 python tools/time_conv.py $SH 2>&1 | grep us; done
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for ms in 8 16; do echo "== step, max split $ms"; OSM_CONV_MAX_SPLIT=$ms timeout 300 python tools/profile_step.py --batch 1 2>&1 | grep -E "^step|\[conv|\[gn" ; done
