#!/bin/bash
mkdir -p gpurun_out
timeout 120 python tools/gpu_mma_bench.py > gpurun_out/mma_bench.log 2>&1; cat gpurun_out/mma_bench.log
