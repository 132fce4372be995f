#!/bin/bash
# ncu --set full captures of the three kernels of a step (run on the GPU box); raw reports stay in gpurun_out/.
TAG=${1:-x}
mkdir -p gpurun_out
BENCH="python bench.py --steps 3 --warmup 3 --no-cpu-baseline"
KERNELS=${2:-"epi_fusion_pipe epi_zgemm epi_stage"}
for k in $KERNELS; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 1 -f -o gpurun_out/prof_${k}_$TAG $BENCH > gpurun_out/prof_${k}_$TAG.log 2>&1
done
ls -la gpurun_out/*.ncu-rep
