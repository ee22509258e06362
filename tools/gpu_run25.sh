#!/bin/bash
mkdir -p gpurun_out
echo "== pair / stream-K parity"; timeout 300 python -m pytest tests/test_kernels_gpu.py -x -q -k "cta_pair or fused_groupnorm" 2>&1 | tail -8
SH="1,256,256,256,256,9 1,256,256,512,256,9 1,256,256,256,512,9 1,128,128,256,256,9 1,128,128,512,512,9 1,128,128,512,256,9 1,64,64,512,512,9 1,64,64,1024,512,9 2,256,256,256,256,9 8,256,256,256,256,9"
for v in 0 1; do echo "== timings OSM_CONV_SK=$v"; OSM_CONV_VERBOSE=1 OSM_CONV_SK=$v timeout 240 python tools/time_conv.py $SH 2>&1 | grep -E " us " ; done
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for v in 0 1; do for b in 1 8; do echo "== step B=$b OSM_CONV_SK=$v"; OSM_CONV_SK=$v timeout 300 python tools/profile_step.py --batch $b 2>&1 | grep -E "^step|\[conv" ; done; done
