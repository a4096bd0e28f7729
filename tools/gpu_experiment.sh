#!/bin/bash
# One forward-performance experiment round on the GPU box (every step under its own timeout):
#   parity tests, then for each environment configuration a per-kernel profile and a short bench.
# usage (through gpurun): bash tools/gpu_experiment.sh "A=1" "FSVC_X=1 FSVC_Y=2" ...
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_grads.py -m gpu -q -x 2>&1 | tail -3
for cfg in "$@"; do
  echo "== $cfg"
  timeout 120 env $cfg FSVC_DEBUG_PLAN=1 python tools/kernel_profile.py > gpurun_out/kp.txt 2>&1
  grep -E "sum_ms" gpurun_out/kp.txt
  grep -E "^  [ls][0-9]\." gpurun_out/kp.txt | awk '{printf "%s=%s ", $1,$2}'; echo
  timeout 200 env $cfg python bench.py --steps 30 --warmup 5 --no-eager 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('ms', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], 'parity', d['parity_max_abs_vs_oracle'], 'launches', d['launches_per_step'])"
  cp gpurun_out/kp.txt "gpurun_out/kp_$(echo $cfg | tr ' =' '__').txt"
done
