#!/bin/bash
# GPU pass 3: GroupNorm statistics fused into conv epilogues
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
timeout 600 python tools/check_variants.py > gpurun_out/variants.log 2>&1; tail -5 gpurun_out/variants.log
timeout 300 python tools/profile_step.py --batch 1 > gpurun_out/step_b1.log 2> gpurun_out/step_b1.err
head -5 gpurun_out/step_b1.log
timeout 300 python tools/profile_step.py --batch 8 > gpurun_out/step_b8.log 2>&1
head -5 gpurun_out/step_b8.log
timeout 600 python bench.py --steps 30 --warmup 3 --no-cpu-baseline > gpurun_out/bench_b1.json 2> gpurun_out/bench_b1.err
tail -2 gpurun_out/bench_b1.json | cut -c1-300
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"gn_" -o gpurun_out/r01_gn_v2 python tools/ncu_misc.py gn > gpurun_out/ncu_gn.log 2>&1
