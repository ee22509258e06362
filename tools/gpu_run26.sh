#!/bin/bash
echo "== pair / stream-K parity"; timeout 300 python -m pytest tests/test_kernels_gpu.py -x -q -k "cta_pair or fused_groupnorm" 2>&1 | tail -3
SH="1,256,256,256,256,9 1,256,256,512,256,9 1,128,128,256,256,9 1,128,128,512,512,9 1,64,64,512,512,9 2,256,256,256,256,9"
for v in 0 1; do echo "== timings OSM_CONV_SK=$v"; OSM_CONV_SK=$v timeout 240 python tools/time_conv.py $SH 2>&1 | grep -E " us " ; done
