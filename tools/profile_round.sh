#!/bin/bash
# Runs on the GPU box (via gpurun): the evidence of a round — GPU test-suite log, bench lines, role timers, the ncu launch list of the
# bench command and one full ncu capture of each kernel of a step.  Outputs land in gpurun_out/; tools/summarize_profiles.py turns
# them into the tracked summaries under profiles/.
TAG=${1:-r2}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -12 > gpurun_out/pytest_gpu_$TAG.txt; tail -3 gpurun_out/pytest_gpu_$TAG.txt
timeout 400 python bench.py --steps 50 --warmup 10 > gpurun_out/bench_${TAG}_n1.json 2> gpurun_out/bench_${TAG}_n1.err
timeout 400 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_${TAG}_reference_arm.json 2>/dev/null
timeout 400 python bench.py --workload cfg3 --steps 50 --warmup 10 --no-cpu-baseline > gpurun_out/bench_${TAG}_cfg3.json 2> gpurun_out/bench_${TAG}_cfg3.err
timeout 600 python bench.py --workload sweep --steps 20 --warmup 5 > gpurun_out/bench_${TAG}_sweep.json 2> gpurun_out/bench_${TAG}_sweep.err
timeout 400 python bench.py --workload cfg4_256 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_${TAG}_cfg4_256.json 2> gpurun_out/bench_${TAG}_cfg4_256.err
timeout 120 python tools/gpu_pipe_timers.py 64 > gpurun_out/pipe_timers_$TAG.txt 2>&1
BENCH="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-gpu-reference"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"epi|split|nchw|z_epi|sector|unstage" -c 60 --csv \
    --log-file gpurun_out/launches_$TAG.csv $BENCH > gpurun_out/launches_$TAG.log 2>&1
bash tools/gpu_prof.sh $TAG
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks.mem,power.draw,clocks_event_reasons.active --format=csv > gpurun_out/smi_$TAG.csv
python - <<PY
import json
for f in ("n1","cfg3","cfg4_256"):
    try:
        d=json.load(open("gpurun_out/bench_${TAG}_%s.json" % f)); print(f, d["ms_per_step"], d["breakdown"], d["e2e"]["ms_per_step"], d.get("gpu_reference",{}).get("ms_per_step"), d.get("cpu_baseline",{}).get("ms_per_step"))
    except Exception as e: print(f, "ERR", e)
PY
ls gpurun_out | tail -30
