#!/bin/bash
mkdir -p gpurun_out
NG=$(nvidia-smi -L | wc -l)
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $NG --steps 100 --warmup 10 > gpurun_out/bench_n${NG}_peer.json 2> gpurun_out/bench_n${NG}_peer.err
python -c "
import json;d=json.loads(open('gpurun_out/bench_n${NG}_peer.json').read().strip().splitlines()[-1]);print('default', d['n_gpus'], round(d['value']), d['ms_per_step'], round(d['e2e']['value']), d['config']['parallelism'][:70])" || (head -c 300 gpurun_out/bench_n${NG}_peer.json; grep -E "Error|error" -B2 -A8 gpurun_out/bench_n${NG}_peer.err | head -40)
wc -l gpurun_out/bench_n${NG}_peer.json
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29542 bench.py --impl reference --gpus $NG --steps 3 --warmup 1 > gpurun_out/bench_n${NG}_ref.json 2> gpurun_out/bench_n${NG}_ref.err
cut -c1-200 gpurun_out/bench_n${NG}_ref.json
