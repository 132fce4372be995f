#!/bin/bash
# Developer loop on the GPU box: variant check, full GPU test-suite, role timers, bench line, ncu launch list.
TAG=${1:-x}
mkdir -p gpurun_out
timeout 200 python tools/gpu_check.py tiny_ring_z tiny_randn_krt cfg1_randn_krt cfg2_r50_256_randn cfg3_r152_384 2>&1 | grep -v sector
if [ "$2" != "notest" ]; then timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -15; fi
timeout 120 python tools/gpu_pipe_timers.py 64 > gpurun_out/pipe_timers_$TAG.txt 2>&1; cat gpurun_out/pipe_timers_$TAG.txt
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-reference > gpurun_out/bench_$TAG.json 2>gpurun_out/bench_$TAG.err; tail -2 gpurun_out/bench_$TAG.err
python -c "
import json; d=json.load(open('gpurun_out/bench_$TAG.json')); print('ms_per_step', d['ms_per_step'], 'breakdown', d['breakdown'], 'e2e', d['e2e']['ms_per_step'], 'gpu_ref', d.get('gpu_reference',{}).get('ms_per_step'))"
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"epi|split|nchw|z_epi|sector" -c 40 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/launches_$TAG.log 2>&1
python - <<PY
import csv,collections
rows=[r for r in csv.reader(open("gpurun_out/launches_$TAG.csv")) if len(r)>5 and r[0].isdigit()]
d=collections.defaultdict(list)
for r in rows: d[r[4]].append(float(r[-1]))
for k,v in d.items(): print(k[:60], len(v), sum(v)/len(v))
PY
