#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -q -x -k "T3 or module or streamer or mpjpe or sweep" > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu.log
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/bench_final_nocpu.json 2> gpurun_out/bench_final_nocpu.err
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"epi|split|nchw|z_epi|sector" -c 60 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/launches_r1.log 2>&1
tail -n 3 gpurun_out/pytest_gpu.log; cat gpurun_out/bench_final_nocpu.json | cut -c1-2400; tail -n 3 gpurun_out/bench_final_nocpu.err
