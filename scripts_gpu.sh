#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu.log
timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/bench_tile.json 2> gpurun_out/bench_tile.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"epi|split|nchw|z_epi" -c 40 --csv --log-file gpurun_out/launches_tile_v6.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
tail -n 6 gpurun_out/pytest_gpu.log; python -c "import json;d=json.load(open('gpurun_out/bench_tile.json'));print(d['ms_per_step'], d['roofline']['kernel_ms'], d['e2e']['ms_per_step'])"; tail -n 3 gpurun_out/bench_tile.err
