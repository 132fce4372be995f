#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 100 --warmup 10 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
cat gpurun_out/gpus.txt; cat gpurun_out/bench_n2.json | cut -c1-1500; tail -n 5 gpurun_out/bench_n2.err
