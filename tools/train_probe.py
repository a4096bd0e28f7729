"""Timing of the native training path (fsvc_forward_train + fsvc_backward) at BASELINE config 3 shape
(B=16, 51 frames = 8160 samples) next to PyTorch eager fwd+bwd of the same op sequence.

    python tools/train_probe.py [--batch 16] [--frames 51]
"""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import harana.models as M
from oracle import fastsvc_torch as otorch
from svcc23_fastsvc_b200 import synthetic as syn

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=16)
ap.add_argument("--frames", type=int, default=51)
ap.add_argument("--no-eager", action="store_true")
ap.add_argument("--once", action="store_true", help="two steps only (what an ncu launch-list pass wraps)")
args = ap.parse_args()
dev = torch.device("cuda:0")
cfg = dict(syn.YAML_CONFIG)
params = syn.make_params(cfg, seed=0, weight_norm=True)
ins = [torch.from_numpy(a).to(dev) for a in syn.make_inputs(args.batch, args.frames, cfg, seed=1)]
g = M.FastSVCGenerator(**{k: (list(v) if isinstance(v, list) else v) for k, v in cfg.items()})
g.load_state_dict({k: torch.from_numpy(v) for k, v in params.items()})
g = g.train().to(dev)
w = torch.randn(args.batch, 1, args.frames * 160, device=dev)


def step():
    g.zero_grad(set_to_none=True)
    y = g(*ins)
    (y * w).sum().backward()


def timed(fn, n=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


if args.once:
    step()
    step()
    torch.cuda.synchronize()
    sys.exit(0)
ms = timed(step)
print(f"native fwd+bwd B={args.batch} frames={args.frames}: {ms:.3f} ms/step, launches fwd+bwd (last call) {g.last_launch_count()}")
with torch.no_grad():
    g.precision = "fp32"
    print(f"fp32 inference forward: {timed(lambda: g(*ins)):.3f} ms")
    g.precision = "auto"
    print(f"auto inference forward: {timed(lambda: g(*ins)):.3f} ms")
if not args.no_eager:
    torch.backends.cudnn.benchmark = True
    tp = {k: torch.from_numpy(v).to(dev).requires_grad_(True) for k, v in params.items()}

    def eager():
        for t in tp.values():
            t.grad = None
        y = otorch.generator_forward(tp, *ins, recompute=True)
        (y * w).sum().backward()

    print(f"torch eager fwd+bwd (reference op sequence, cuDNN): {timed(eager, 5, 2):.3f} ms/step")
