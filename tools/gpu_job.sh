#!/bin/bash
# One gpurun job: $1 selects what to run (scratch helper; outputs under gpurun_out/).
mkdir -p gpurun_out
case "$1" in
  tests)
    python -m pytest tests -m gpu -x -q ${2:+-k "$2"} -s 2>&1 | tail -150 > gpurun_out/pytest_gpu.log; tail -5 gpurun_out/pytest_gpu.log ;;
  bench)
    python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -c 600 gpurun_out/bench_n1.err; head -c 3000 gpurun_out/bench_n1.json ;;
  refarms)
    python bench.py --impl reference-gpu --steps 20 --warmup 5 > gpurun_out/bench_refgpu.json 2> gpurun_out/bench_refgpu.err; cat gpurun_out/bench_refgpu.json
    python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_refcpu.json 2> gpurun_out/bench_refcpu.err; cat gpurun_out/bench_refcpu.json ;;
  halotime)
    for shp in "1 256 256 256 256" "1 256 256 512 256" "4 256 256 256 256" "1 128 128 256 256" "1 128 128 512 512"; do for tn in 256 128; do for xf in 0 1; do timeout 300 python tools/halo_probe.py $tn $xf $shp --time 2>&1 | tail -2 | tr '
' ' '; echo; done; done; done ;;
esac
