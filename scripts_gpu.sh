#!/bin/bash
mkdir -p gpurun_out
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu.log
timeout 400 python bench.py > gpurun_out/bench_final_n1.json 2> gpurun_out/bench_final_n1.err
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_final_ref.json 2> gpurun_out/bench_final_ref.err
timeout 300 python bench.py --workload cfg3 --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/bench_final_cfg3.json 2> gpurun_out/bench_final_cfg3.err
./tools/profile_round.sh r1 > gpurun_out/profile_round.log 2>&1
tail -n 3 gpurun_out/smoke.log; tail -n 3 gpurun_out/pytest_gpu.log; wc -l gpurun_out/bench_final_n1.json; cat gpurun_out/bench_final_n1.json | cut -c1-3000; cut -c1-400 gpurun_out/bench_final_ref.json; python -c "import json;d=json.load(open('gpurun_out/bench_final_cfg3.json'));print('cfg3', d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'])"
