#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
for v in 0 1; do for b in 1 8; do echo "== step B=$b OSM_GN_SMALL=$v"; OSM_GN_SMALL=$v timeout 300 python tools/profile_step.py --batch $b 2>&1 | grep -E "^step|^launches|\[gn" ; done; done
