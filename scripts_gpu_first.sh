#!/bin/bash
# first GPU visit: smoke, parity tests, bench (both arms short), launch list
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
nproc > gpurun_out/nproc.txt
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 300 python bench.py --steps 100 --warmup 10 --cpu-steps 5 > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err
timeout 300 python bench.py --workload cfg3 --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/bench_cfg3.json 2> gpurun_out/bench_cfg3.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_r1_v0.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
tail -5 gpurun_out/smoke.log gpurun_out/pytest_gpu.log; cat gpurun_out/bench_ours.json gpurun_out/bench_cfg3.json; tail -3 gpurun_out/bench_ours.err
