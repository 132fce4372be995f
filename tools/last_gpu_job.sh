#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/gpu_sweep.py > gpurun_out/sweep.log 2>&1; tail -n 22 gpurun_out/sweep.log
