#!/bin/bash
# Multi-GPU evidence on the box (gpurun --gpus N): 2-GPU tests + bench lines per exchange mode.
N=${1:-2}; TAG=${2:-r2}
mkdir -p gpurun_out
if [ "$N" = "2" ]; then timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q 2>&1 | tail -4 | tee gpurun_out/pytest_gpu_multi_$TAG.txt; fi
PORT=29511
for X in ${MODES:-peer p2p allgather}; do
  PORT=$((PORT+1))
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $PORT bench.py --gpus $N --steps 50 --warmup 10 --exchange $X --no-cpu-baseline --no-gpu-reference ${3:-} > gpurun_out/bench_${TAG}_n${N}_$X.json 2> gpurun_out/bench_${TAG}_n${N}_$X.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_${TAG}_n${N}_$X.json").read().strip().splitlines()[-1])
    print("$X", "ms_per_step", d["ms_per_step"], "value", d["value"], "e2e", d["e2e"]["ms_per_step"], "parity", d.get("config",{}).get("parity") or d.get("parity"), d.get("breakdown"))
except Exception as e:
    print("$X ERR", e); print(open("gpurun_out/bench_${TAG}_n${N}_$X.err").read()[-1500:])
PY
done
