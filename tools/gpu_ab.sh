#!/bin/bash
# A/B on one box: alternate configurations several times, short benches (ms_per_step only).
# usage: bash tools/gpu_ab.sh ROUNDS "A=1" "FSVC_X=1" ...
rounds=$1; shift
for r in $(seq 1 $rounds); do
  for cfg in "$@"; do
    echo -n "$cfg : "
    timeout 200 env $cfg python bench.py --steps 40 --warmup 8 --no-eager --quick 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('ms', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['ms_per_step'],4))"
  done
done
