"""Per-launch device times of one generator forward at BASELINE config 2 (B=32, 100 frames).

    python tools/kernel_profile.py [--mode auto|fp32|tc_bf16x3] [--out gpurun_out/kernels.json] [--once]

--once runs a single warm forward + a few timed ones (what the ncu passes wrap).
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import harana.models as M
from svcc23_fastsvc_b200 import synthetic as syn

ap = argparse.ArgumentParser()
ap.add_argument("--mode", default="auto")
ap.add_argument("--out", default="")
ap.add_argument("--once", action="store_true")
ap.add_argument("--batch", type=int, default=32)
ap.add_argument("--frames", type=int, default=100)
args = ap.parse_args()

dev = torch.device("cuda:0")
cfg = dict(syn.YAML_CONFIG)
params = syn.make_params(cfg, seed=0)
ins = syn.make_inputs(args.batch, args.frames, cfg, seed=1234)
g = M.FastSVCGenerator(**{k: (list(v) if isinstance(v, list) else v) for k, v in cfg.items()})
g.remove_weight_norm()
g.load_state_dict({k: torch.from_numpy(v) for k, v in params.items()})
g = g.eval().to(dev)
g.precision = args.mode
devin = [torch.from_numpy(a).to(dev) for a in ins]
with torch.no_grad():
    if args.once:
        for _ in range(3):
            g(*devin)
        torch.cuda.synchronize()
        sys.exit(0)
    for _ in range(3):
        g(*devin)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    runs = []
    for _ in range(5):
        flush.zero_()
        runs.append(g.profile(*devin))
recs = runs[-1]
for i, r in enumerate(recs):
    r["ms"] = sorted(run[i]["ms"] for run in runs)[len(runs) // 2]
total = sum(r["ms"] for r in recs)
print(f"mode={args.mode} launches={len(recs)} sum_ms={total:.3f}")
for r in recs:
    gbs = r["bytes"] / (r["ms"] * 1e-3) / 1e9 if r["ms"] > 0 else 0
    tf = r["flops"] / (r["ms"] * 1e-3) / 1e12 if r["ms"] > 0 else 0
    print(f"  {r['label']:24s} {r['ms']*1e3:8.1f} us  {100*r['ms']/total:5.1f}%  {gbs:7.0f} GB/s  {tf:6.1f} TF/s")
if args.out:
    with open(args.out, "w") as f:
        json.dump({"mode": args.mode, "sum_ms": total, "records": recs}, f, indent=1)
