#!/bin/bash
# Multi-GPU bench lines (one box, N GPUs): bash tools/gpu_multi.sh N [configs...]
#   -> gpurun_out/r2_{bench,train,convert}_${N}gpu.json   (configs default: infer train convert)
N=$1; shift
CFGS=${@:-infer train convert}
run() { timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 200)) bench.py --gpus $N "$@"; }
for c in $CFGS; do
  case $c in
    infer) run --steps 30 --warmup 5 --no-eager 2>/dev/null | grep '^{' | tail -1 > gpurun_out/r2_bench_${N}gpu.json ;;
    train) run --config train 2>/dev/null | grep '^{' | tail -1 > gpurun_out/r2_train_${N}gpu.json ;;
    convert) run --config convert 2>/dev/null | grep '^{' | tail -1 > gpurun_out/r2_convert_${N}gpu.json ;;
  esac
done
python - <<PY
import json, os
for k in ("bench","train","convert"):
    f=f"gpurun_out/r2_{k}_${N}gpu.json"
    if not os.path.exists(f): continue
    try:
        d=json.loads(open(f).read()); print(k, d["n_gpus"], d["value"], d["unit"], d.get("ms_per_step"), d.get("wall_clock_s"), d.get("e2e",{}).get("value"))
    except Exception as e: print(k, "failed", e)
PY
