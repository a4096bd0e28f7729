"""How much of the forward is launch gap: the same forward replayed from a CUDA graph vs launched kernel by kernel."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import harana.models as M
from svcc23_fastsvc_b200 import synthetic as syn

dev = torch.device("cuda:0")
cfg = dict(syn.YAML_CONFIG)
params = syn.make_params(cfg, seed=0)
g = M.FastSVCGenerator(**{k: (list(v) if isinstance(v, list) else v) for k, v in cfg.items()})
g.remove_weight_norm()
g.load_state_dict({k: torch.from_numpy(v) for k, v in params.items()})
g = g.eval().to(dev)
devin = [torch.from_numpy(a).to(dev) for a in syn.make_inputs(32, 100, cfg, seed=1234)]
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

def timed(fn, n=30):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(n):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / n

with torch.no_grad():
    print("stream launches: %.4f ms" % timed(lambda: g(*devin)))
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(3):
            y = g(*devin)
    torch.cuda.current_stream().wait_stream(s)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        y = g(*devin)
    ref = g(*devin)
    graph.replay()
    torch.cuda.synchronize()
    print("graph == stream result:", bool(torch.equal(y, ref)))
    print("graph replay:    %.4f ms" % timed(graph.replay))
