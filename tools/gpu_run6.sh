#!/bin/bash
# GPU pass 6: programmatic dependent launch on/off
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -8 gpurun_out/pytest_gpu.log
timeout 300 python tools/check_variants.py > gpurun_out/variants.log 2>&1; tail -5 gpurun_out/variants.log
for pdl in 1 0; do
  OSM_PDL=$pdl timeout 300 python tools/profile_step.py --batch 1 > gpurun_out/step_b1_pdl$pdl.log 2> gpurun_out/step_b1.err
  echo "PDL=$pdl B=1"; head -2 gpurun_out/step_b1_pdl$pdl.log
  OSM_PDL=$pdl timeout 300 python tools/profile_step.py --batch 8 > gpurun_out/step_b8_pdl$pdl.log 2>&1
  echo "PDL=$pdl B=8"; grep "^step" gpurun_out/step_b8_pdl$pdl.log
done
