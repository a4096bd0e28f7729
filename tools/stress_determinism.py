"""Run the config-2 forward repeatedly and compare every run bit-for-bit with the first (race detector)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import harana.models as M
from svcc23_fastsvc_b200 import synthetic as syn
n = int(sys.argv[1]) if len(sys.argv) > 1 else 30
cfg = dict(syn.YAML_CONFIG); params = syn.make_params(cfg, seed=0)
ins = [torch.from_numpy(a).cuda() for a in syn.make_inputs(32, 100, cfg, seed=1234)]
g = M.FastSVCGenerator(**{k: (list(v) if isinstance(v, list) else v) for k, v in cfg.items()})
g.remove_weight_norm(); g.load_state_dict({k: torch.from_numpy(v) for k, v in params.items()}); g = g.eval().cuda()
with torch.no_grad():
    g.precision = "fp32"; ref = g(*ins).clone()
    g.precision = "auto"; y0 = g(*ins).clone()
    print("auto vs fp32 max-abs", float((y0 - ref).abs().max()))
    bad = 0
    for i in range(n):
        y = g(*ins)
        if not torch.equal(y, y0):
            bad += 1
            d = (y - y0).abs()
            print("run", i, "differs: max", float(d.max()), "count", int((d > 0).sum()), "b", int(d.amax(dim=(1, 2)).argmax()))
    print("nondeterministic runs:", bad, "of", n)
