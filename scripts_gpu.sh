#!/bin/bash
mkdir -p gpurun_out
NG=$(nvidia-smi -L | wc -l)
for EX in peer allgather; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $NG --steps 100 --warmup 10 --exchange $EX > gpurun_out/bench_n${NG}_${EX}.json 2> gpurun_out/bench_n${NG}_${EX}.err
python -c "
import json;d=json.loads(open('gpurun_out/bench_n${NG}_${EX}.json').read().strip().splitlines()[-1]);print('$EX', d['n_gpus'], round(d['value']), d['ms_per_step'], round(d['e2e']['value']), d['config']['parallelism'][:60])" || (head -c 300 gpurun_out/bench_n${NG}_${EX}.json; grep -E "Error|error" -B2 -A8 gpurun_out/bench_n${NG}_${EX}.err | head -40)
done
