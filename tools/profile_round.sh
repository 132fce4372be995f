#!/bin/bash
# Runs on the GPU box (via gpurun): launch list of the bench command + one full ncu capture of each hot kernel.
# Outputs land in gpurun_out/; summaries are copied into profiles/ by tools/summarize_profiles.py.
TAG=${1:-r1}
mkdir -p gpurun_out
BENCH="python bench.py --steps 3 --warmup 3 --no-cpu-baseline"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"epi|split|nchw|z_epi|sector" -c 60 --csv \
    --log-file gpurun_out/launches_${TAG}.csv $BENCH > gpurun_out/launches_${TAG}.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:epi_fusion_tile -s 2 -c 1 -f \
    -o gpurun_out/prof_tile_${TAG} $BENCH > gpurun_out/prof_tile_${TAG}.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:epi_zgemm -s 2 -c 1 -f \
    -o gpurun_out/prof_zgemm_${TAG} $BENCH > gpurun_out/prof_zgemm_${TAG}.log 2>&1
timeout 900 ncu --set full --clock-control none -k regex:split_planes -s 2 -c 1 -f \
    -o gpurun_out/prof_split_${TAG} $BENCH > gpurun_out/prof_split_${TAG}.log 2>&1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks.mem,power.draw,clocks_event_reasons.active --format=csv > gpurun_out/smi_${TAG}.csv
ls -la gpurun_out | tail -12
