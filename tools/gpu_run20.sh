#!/bin/bash
for m in 1 2048 4096 8192 16384 32768; do for b in 1 2; do echo "== B=$b OSM_GN_SMALL_MAX=$m"; OSM_GN_SMALL_MAX=$m timeout 300 python tools/profile_step.py --batch $b 2>&1 | grep -E "^step|^launches" ; done; done
