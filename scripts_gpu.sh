#!/bin/bash
mkdir -p gpurun_out
timeout 120 python tools/gpu_check.py tiny_ring_z cfg1_ring cfg2_r50_256_randn cfg3_r152_384 > gpurun_out/check.log 2>&1; echo "rc=$?" >> gpurun_out/check.log
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu.log
timeout 300 python tools/gpu_timers.py 64 > gpurun_out/timers.log 2>&1
timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/bench_tile.json 2> gpurun_out/bench_tile.err
timeout 120 python tools/gpu_mma_bench.py > gpurun_out/mma_bench.log 2>&1
cat gpurun_out/check.log; tail -n 4 gpurun_out/pytest_gpu.log; head -12 gpurun_out/timers.log; python -c "import json;d=json.load(open('gpurun_out/bench_tile.json'));print(d['ms_per_step'], d['roofline']['kernel_ms'], d['e2e']['ms_per_step'])"; tail -n 3 gpurun_out/bench_tile.err; cat gpurun_out/mma_bench.log
