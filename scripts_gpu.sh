#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "streamer" > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu.log
timeout 400 python bench.py --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err
tail -n 4 gpurun_out/pytest_gpu.log; python -c "import json;d=json.load(open('gpurun_out/bench_ours.json'));print(d['ms_per_step'], d['e2e'])"; tail -n 3 gpurun_out/bench_ours.err
