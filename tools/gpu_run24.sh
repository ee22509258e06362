#!/bin/bash
SH="1,256,256,256,256,9 1,256,256,512,256,9 1,256,256,256,512,9 1,128,128,256,256,9 1,128,128,512,512,9 1,128,128,512,256,9 1,64,64,512,512,9 2,256,256,256,256,9"
T="timeout 200 python tools/time_conv.py"
echo "== default policy";  $T $SH 2>&1 | grep us
echo "== forced 128,1 (single-CTA persistent, BN=128)"; OSM_CONV_FORCE=128,1 $T $SH 2>&1 | grep us
echo "== forced 256,1 2SM off"; OSM_CONV_2SM=0 OSM_CONV_FORCE=256,1 $T $SH 2>&1 | grep us
echo "== forced 256,1 2SM forced"; OSM_CONV_2SM=2 OSM_CONV_FORCE=256,1 $T $SH 2>&1 | grep us
echo "== forced 64,1"; OSM_CONV_FORCE=64,1 $T 1,256,256,256,256,9 1,128,128,256,256,9 1,64,64,512,512,9 2>&1 | grep us
