#!/bin/bash
# One gpurun job: $1 selects what to run (scratch helper; outputs under gpurun_out/).
mkdir -p gpurun_out
case "$1" in
  tests)
    timeout 1500 python -m pytest tests -m gpu -x -q ${2:+-k "$2"} -s 2>&1 | tail -150 > gpurun_out/pytest_gpu.log; tail -5 gpurun_out/pytest_gpu.log ;;
  bench)
    timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -c 600 gpurun_out/bench_n1.err; head -c 3000 gpurun_out/bench_n1.json ;;
  refarms)
    python bench.py --impl reference-gpu --steps 20 --warmup 5 > gpurun_out/bench_refgpu.json 2> gpurun_out/bench_refgpu.err; cat gpurun_out/bench_refgpu.json
    python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_refcpu.json 2> gpurun_out/bench_refcpu.err; cat gpurun_out/bench_refcpu.json ;;
  h16)
    for mode in 0 1 2; do timeout 120 python tools/halo16_probe.py $mode 2 32 32 64 256 2>&1 | tail -3; done
    timeout 120 python tools/halo16_probe.py 2 1 64 64 256 256 2>&1 | tail -3
    timeout 120 python tools/halo16_probe.py 2 3 16 8 128 512 2>&1 | tail -3
    for shp in "1 256 256 256 256" "1 256 256 512 256" "4 256 256 256 256" "1 128 128 256 256" "1 64 64 512 512" "8 64 64 512 512" "8 32 32 512 512"; do for mode in 0 2; do timeout 300 python tools/halo16_probe.py $mode $shp --time 2>&1 | tail -2 | tr '\n' ' '; echo; done; done ;;
  prof)
    timeout 600 python tools/profile_step.py --batch 1 > gpurun_out/step_breakdown_b1.log 2>&1
    timeout 900 python tools/profile_step.py --batch ${2:-32} > gpurun_out/step_breakdown_b${2:-32}.log 2>&1 ;;
  ncu16)
    timeout 600 ncu --set full --clock-control none --import-source on -k regex:halo16 -s 2 -c 1 -f -o gpurun_out/ncu_halo16 python tools/halo16_probe.py ${2:-2} ${3:-4 256 256 256 256} --time > gpurun_out/ncu16.log 2>&1; tail -3 gpurun_out/ncu16.log ;;
  launches)
    # ncu launch list of ONE steady-state region of the bench (cold-cache, serialised per-launch times: compare shares)
    timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches.csv \
        python bench.py --steps 2 --warmup 3 --no-table --no-cpu-baseline --no-eager-reference --profiler-range > gpurun_out/launches_bench.log 2>&1
    tail -2 gpurun_out/launches_bench.log | cut -c1-300; wc -l gpurun_out/launches.csv ;;
  ncu16s)
    for m in 1 2; do timeout 120 python tools/halo16_stats_probe.py $m 4 256 256 256 256 2>&1 | tail -1; done
    timeout 600 ncu --set full --clock-control none --import-source on -k regex:halo16 -s 2 -c 1 -f -o gpurun_out/ncu_halo16s python tools/halo16_stats_probe.py 2 4 256 256 256 256 > gpurun_out/ncu16s.log 2>&1; tail -2 gpurun_out/ncu16s.log ;;
esac
