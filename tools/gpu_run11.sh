#!/bin/bash
# measurement pass for profiles/: bench lines, launch list, ncu --set full of the three kernel families
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/bench_b1.json 2> gpurun_out/bench_b1.err; tail -c 600 gpurun_out/bench_b1.json
timeout 600 python bench.py --batch 8 --no-cpu-baseline --steps 30 --warmup 3 > gpurun_out/bench_b8.json 2> gpurun_out/bench_b8.err; tail -c 300 gpurun_out/bench_b8.json
timeout 600 python bench.py --batch 32 --no-cpu-baseline --steps 10 --warmup 3 > gpurun_out/bench_b32.json 2> gpurun_out/bench_b32.err; tail -c 300 gpurun_out/bench_b32.json
timeout 400 python bench.py --impl reference --steps 4 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; tail -c 400 gpurun_out/bench_ref.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-cuda-graph > gpurun_out/launches_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_tc -o gpurun_out/r01_conv_final python tools/ncu_conv.py 1,256,256,256,256,9 8,256,256,256,256,9 1,8,8,1024,1024,9 > gpurun_out/ncu_conv.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"gn_|flash_|tok_to" -o gpurun_out/r01_gn_flash_final python tools/ncu_misc.py all 1 > gpurun_out/ncu_misc.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"gn_|flash_" -o gpurun_out/r01_gn_flash_b8 python tools/ncu_misc.py all 8 > gpurun_out/ncu_misc8.log 2>&1
ls -la gpurun_out | tail -20
