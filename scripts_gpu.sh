#!/bin/bash
mkdir -p gpurun_out
timeout 600 compute-sanitizer --tool synccheck --error-exitcode 9 python tools/gpu_check.py tiny_ring_z tiny_randn_krt > gpurun_out/sanitizer_synccheck.log 2>&1; echo "synccheck rc=$?" >> gpurun_out/sanitizer_synccheck.log
timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/bench_tile.json 2> gpurun_out/bench_tile.err
tail -n 5 gpurun_out/sanitizer_synccheck.log | cut -c1-200; python -c "import json;d=json.load(open('gpurun_out/bench_tile.json'));print(d['ms_per_step'], d['roofline']['kernel_ms'], d['e2e']['ms_per_step'])"
