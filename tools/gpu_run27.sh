#!/bin/bash
SH="1,256,256,256,256,9 8,256,256,256,256,9 1,256,256,512,256,9 8,128,128,512,512,9 2,256,256,256,256,9,1"
for v in 6 7; do echo "== OSM_CONV_2SM_STAGES=$v"; OSM_CONV_2SM_STAGES=$v timeout 240 python tools/time_conv.py $SH 2>&1 | grep " us "; done
for v in 4 8; do echo "== OSM_CONV_EPI_WARPS=$v"; OSM_CONV_EPI_WARPS=$v timeout 240 python tools/time_conv.py $SH 2>&1 | grep " us "; done
