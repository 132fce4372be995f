#!/bin/bash
# usage: scripts_gpu.sh <tag> [steps...]   — GPU visit helper; everything lands in gpurun_out/
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_umma.py -x -q > gpurun_out/pytest_umma.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_umma.log
timeout 900 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_umma.py > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:epi -c 40 --csv --log-file gpurun_out/launches_r1_v0.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
tail -n 25 gpurun_out/pytest_umma.log; tail -n 8 gpurun_out/pytest_gpu.log
