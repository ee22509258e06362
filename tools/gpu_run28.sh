#!/bin/bash
# end-of-round measurement refresh: tests, bench lines, per-op breakdown, steady-state launch list, GroupNorm ncu capture
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -2 gpurun_out/pytest_gpu.log
timeout 900 python bench.py > gpurun_out/bench_b1.json 2> gpurun_out/bench_b1.err; head -c 400 gpurun_out/bench_b1.json; echo
timeout 600 python bench.py --batch 8 --no-cpu-baseline --steps 30 --warmup 3 > gpurun_out/bench_b8.json 2> gpurun_out/bench_b8.err; head -c 300 gpurun_out/bench_b8.json; echo
timeout 600 python bench.py --batch 32 --no-cpu-baseline --steps 10 --warmup 3 > gpurun_out/bench_b32.json 2> gpurun_out/bench_b32.err; head -c 300 gpurun_out/bench_b32.json; echo
timeout 400 python bench.py --impl reference --steps 4 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; head -c 300 gpurun_out/bench_ref.json; echo
for b in 1 8; do timeout 300 python tools/profile_step.py --batch $b > gpurun_out/step_b$b.log 2>&1; done
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-cuda-graph --profiler-range > gpurun_out/launches_bench.log 2>&1
python tools/summarize_launches.py gpurun_out/launches.csv > gpurun_out/launches_summary.md
for b in 1 8; do
  timeout 600 ncu --set full --clock-control none -k regex:"gn_" -o /tmp/gn_b$b python tools/ncu_misc.py gn $b > gpurun_out/ncu_misc_b$b.log 2>&1
  python tools/ncu_summary.py /tmp/gn_b$b.ncu-rep > gpurun_out/ncu_gn_b${b}_summary.md
done
du -sh gpurun_out
