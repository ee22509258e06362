#!/bin/bash
for v in 0 1; do for b in 1 8; do echo "== B=$b OSM_PDL=$v"; OSM_PDL=$v timeout 300 python tools/profile_step.py --batch $b 2>&1 | grep -E "^step" ; done; done
OSM_PDL=1 timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
