"""GPU probe: compare fp32-mode and tensor-core-mode results block by block (debug aid)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from harana.models.fastsvc import FastSVCDownsampleNet, FastSVCFiLMNet, FastSVCUpsampleNet
import harana.models as M
from svcc23_fastsvc_b200 import synthetic as syn

dev = torch.device("cuda:0")
torch.manual_seed(0)

def cmp(name, a, b):
    a, b = a.float().cpu(), b.float().cpu()
    d = (a - b).abs()
    print(f"{name:28s} max-abs {d.max().item():.3e}  mean-abs {d.mean().item():.3e}  ref-absmax {b.abs().max().item():.3f}"
          f"  argmax {np.unravel_index(int(d.argmax()), d.shape)}", flush=True)

for C, T in ((24, 300), (48, 200), (96, 130), (192, 70)):
    f = FastSVCFiLMNet(C).to(dev)
    x = torch.randn(2, C, T, device=dev)
    with torch.no_grad():
        f.precision = "fp32"; s0, h0 = f(x)
        f.precision = "tc_bf16x3"; s1, h1 = f(x)
    torch.cuda.synchronize()
    cmp(f"film C={C} scale", s1, s0); cmp(f"film C={C} shift", h1, h0)

for cin, c, sc, T in ((24, 48, 5, 800), (48, 96, 4, 640), (96, 192, 4, 320)):
    d = FastSVCDownsampleNet(cin, c, sc).to(dev)
    x = torch.randn(2, cin, T, device=dev)
    with torch.no_grad():
        d.precision = "fp32"; y0 = d(x)
        d.precision = "tc_bf16x3"; y1 = d(x)
    cmp(f"down {cin}->{c}/{sc}", y1, y0)

for cin, c, r, T in ((48, 24, 5, 160), (144, 192, 2, 50), (96, 48, 4, 100)):
    u = FastSVCUpsampleNet(cin, c, r, 512, True).to(dev)
    x = torch.randn(2, cin, T, device=dev)
    gb = [torch.randn(2, c, T * r, device=dev) for _ in range(4)]
    spk = torch.randn(2, 512, device=dev)
    with torch.no_grad():
        for s in (spk, None):
            u.precision = "fp32"; y0 = u(x, (gb[0], gb[1]), (gb[2], gb[3]), s)
            u.precision = "tc_bf16x3"; y1 = u(x, (gb[0], gb[1]), (gb[2], gb[3]), s)
            cmp(f"up {cin}->{c}x{r} spk={s is not None}", y1, y0)

cfg = dict(syn.YAML_CONFIG)
params = syn.make_params(cfg, seed=0)
ppg, sine, lft, spk = syn.make_inputs(4, 20, cfg, seed=5)
g = M.FastSVCGenerator(**{k: (list(v) if isinstance(v, list) else v) for k, v in cfg.items()})
g.remove_weight_norm(); g.load_state_dict({k: torch.from_numpy(v) for k, v in params.items()}); g = g.eval().to(dev)
ins = [torch.from_numpy(a).to(dev) for a in (ppg, sine, lft, spk)]
with torch.no_grad():
    g.precision = "fp32"; y0 = g(*ins)
    g.precision = "tc_bf16x3"; y1 = g(*ins)
cmp("generator B=4 f=20", y1, y0)
print("launches", g.last_launch_count())
