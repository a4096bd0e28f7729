#!/bin/bash
# per-kernel profile under each environment configuration: bash tools/gpu_kp.sh "A=1" "FSVC_X=1" ...
for cfg in "$@"; do
  echo "== $cfg"
  timeout 120 env $cfg python tools/kernel_profile.py > gpurun_out/kp.txt 2>&1
  grep -E "sum_ms" gpurun_out/kp.txt
  grep -E "^  [ls][0-9]\." gpurun_out/kp.txt | awk '{printf "%s=%s ", $1,$2}'; echo
  timeout 200 env $cfg python bench.py --steps 40 --warmup 8 --no-eager --quick 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('ms', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['ms_per_step'],4))"
done
