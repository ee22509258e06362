#!/bin/bash
# CTA-pair (cta_group::2) conv kernel: correctness (forced wherever the plan is 256-wide persistent), then timings against the 1-CTA kernel
mkdir -p gpurun_out
echo "== parity, OSM_CONV_2SM=2 (forced)"; OSM_CONV_2SM=2 OSM_CONV_NO_SPLIT=1 timeout 180 python -m pytest tests/test_kernels_gpu.py -x -q -k "conv" 2>&1 | tail -15
SH="1,256,256,256,256,9 8,256,256,256,256,9 1,256,256,512,256,9 1,128,128,256,256,9 8,128,128,512,512,9 1,256,256,512,256,1 1,64,64,512,512,9 2,256,256,256,256,9,1"
for v in 0 1 2; do echo "== timings OSM_CONV_2SM=$v"; OSM_CONV_2SM=$v timeout 240 python tools/time_conv.py $SH 2>&1 | grep us; done
