#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
timeout 400 python bench.py --cpu-steps 4 > gpurun_out/bench_final_n1.json 2> gpurun_out/bench_final_n1.err
./tools/profile_round.sh r1 > gpurun_out/profile_round.log 2>&1
tail -n 4 gpurun_out/pytest_gpu.log; tail -n 2 gpurun_out/smoke.log; cat gpurun_out/bench_final_n1.json | cut -c1-2600
