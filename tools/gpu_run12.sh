#!/bin/bash
# measurement pass for profiles/: gpu tests, bench lines, launch list, ncu --set full of the three kernel families.
# .ncu-rep files are summarised ON the box (raw page csv + markdown) and only the conv one is kept: gpurun_out/ is capped at 64 MiB.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
timeout 900 python bench.py > gpurun_out/bench_b1.json 2> gpurun_out/bench_b1.err; head -c 900 gpurun_out/bench_b1.json; echo
timeout 600 python bench.py --batch 8 --no-cpu-baseline --steps 30 --warmup 3 > gpurun_out/bench_b8.json 2> gpurun_out/bench_b8.err; head -c 300 gpurun_out/bench_b8.json; echo
timeout 600 python bench.py --batch 32 --no-cpu-baseline --steps 10 --warmup 3 > gpurun_out/bench_b32.json 2> gpurun_out/bench_b32.err; head -c 300 gpurun_out/bench_b32.json; echo
timeout 400 python bench.py --impl reference --steps 4 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; head -c 300 gpurun_out/bench_ref.json; echo
for b in 1 8; do timeout 300 python tools/profile_step.py --batch $b > gpurun_out/step_b$b.log 2>&1; done
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-cuda-graph --profiler-range > gpurun_out/launches_bench.log 2>&1
python tools/summarize_launches.py gpurun_out/launches.csv > gpurun_out/launches_summary.md
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_tc -o gpurun_out/r01_conv_final python tools/ncu_conv.py 1,256,256,256,256,9 8,256,256,256,256,9 1,8,8,1024,1024,9 > gpurun_out/ncu_conv.log 2>&1
python tools/ncu_summary.py gpurun_out/r01_conv_final.ncu-rep > gpurun_out/ncu_conv_summary.md
ncu -i gpurun_out/r01_conv_final.ncu-rep --page raw --csv > gpurun_out/ncu_conv_raw.csv 2>/dev/null
for b in 1 8; do
  timeout 600 ncu --set full --clock-control none -k regex:"gn_|flash_|tok_to" -o /tmp/gnf_b$b python tools/ncu_misc.py all $b > gpurun_out/ncu_misc_b$b.log 2>&1
  python tools/ncu_summary.py /tmp/gnf_b$b.ncu-rep > gpurun_out/ncu_gn_flash_b${b}_summary.md
  ncu -i /tmp/gnf_b$b.ncu-rep --page raw --csv > gpurun_out/ncu_gn_flash_b${b}_raw.csv 2>/dev/null
done
timeout 300 ncu --set full --clock-control none -k regex:"guidance_phi|sampler_update|posterior" -c 12 -o /tmp/samp python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-cuda-graph > gpurun_out/ncu_sampler.log 2>&1
python tools/ncu_summary.py /tmp/samp.ncu-rep > gpurun_out/ncu_sampler_summary.md
du -sh gpurun_out; ls -la gpurun_out | tail -30
