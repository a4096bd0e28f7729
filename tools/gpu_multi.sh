#!/bin/bash
# Multi-GPU bench lines (one box, N GPUs): bash tools/gpu_multi.sh N  -> gpurun_out/r2_{bench,train,convert}_${N}gpu.json
N=$1
run() { timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 200)) bench.py --gpus $N "$@"; }
run --steps 30 --warmup 5 --no-eager 2>/dev/null | grep '^{' | tail -1 > gpurun_out/r2_bench_${N}gpu.json
run --config train 2>/dev/null | grep '^{' | tail -1 > gpurun_out/r2_train_${N}gpu.json
run --config convert 2>/dev/null | grep '^{' | tail -1 > gpurun_out/r2_convert_${N}gpu.json
python - <<PY
import json
for k in ("bench","train","convert"):
    try:
        d=json.loads(open(f"gpurun_out/r2_{k}_${N}gpu.json").read()); print(k, d["n_gpus"], d["value"], d["unit"], d.get("ms_per_step"), d.get("wall_clock_s"), d.get("e2e",{}).get("value"))
    except Exception as e: print(k, "failed", e)
PY
