"""Event timeline of the conv_tc3 launches of one forward (needs a libfsvc.so built with -DFSVC_TIMELINE):

    nvcc ... -DFSVC_TIMELINE -o svcc23_fastsvc_b200/libfsvc.so svcc23_fastsvc_b200/csrc/fsvc_abi.cu
    python tools/timeline.py

Prints, per launch (CTA 0), microseconds since kernel entry of: setup done, transform past griddepcontrol.wait, first
rows landed, A blocks converted, MMA saw A blocks, last commit, epilogue saw the accumulator, epilogue done, exit."""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import harana.models as M
from svcc23_fastsvc_b200 import abi, synthetic as syn

dev = torch.device("cuda:0")
cfg = dict(syn.YAML_CONFIG)
params = syn.make_params(cfg, seed=0)
ins = syn.make_inputs(32, 100, cfg, seed=1234)
g = M.FastSVCGenerator(**{k: (list(v) if isinstance(v, list) else v) for k, v in cfg.items()})
g.remove_weight_norm()
g.load_state_dict({k: torch.from_numpy(v) for k, v in params.items()})
g = g.eval().to(dev)
devin = [torch.from_numpy(a).to(dev) for a in ins]
with torch.no_grad():
    for _ in range(3):
        g(*devin)
    torch.cuda.synchronize()
    recs = g.profile(*devin)     # serialised launches: each kernel's timeline is its own
torch.cuda.synchronize()
n = 64 * 8 * 64
buf = np.zeros(n, dtype=np.uint64)
lib = abi.load()
lib.fsvc_debug_timeline.argtypes = [ctypes.c_void_p, ctypes.c_int]
abi.check(lib.fsvc_debug_timeline(buf.ctypes.data, n))
tl = buf.reshape(64, 8, 64).astype(np.int64)
labels = [r["label"] for r in recs]
print(f"{'launch':22s} {'ms':>6s} | setup  gdw   rows | A blocks converted ... | MMA saw A ... | commit accF  epiE  exit")
for slot, lab in enumerate(labels[:64]):
    t = tl[slot, 0]
    if t[0] == 0 or t[40] == 0:
        continue
    us = lambda e: (t[e] - t[0]) / 1e3 if t[e] else float("nan")
    ab = " ".join(f"{us(4 + i):5.1f}" for i in range(12) if t[4 + i])
    mb = " ".join(f"{us(20 + i):5.1f}" for i in range(12) if t[20 + i])
    print(f"{lab:22s} {recs[slot]['ms']*1e3:6.1f} | {us(1):5.1f} {us(2):5.1f} {us(3):5.1f} | {ab} | {mb} | {us(36):5.1f} {us(38):5.1f} {us(39):5.1f} {us(40):5.1f}")
    if t[41] and t[47]:   # steady state, item 5: inside one iteration of each role (us since kernel entry)
        print(f"    item 5  transform: top {us(41):.2f} staged {us(42):.2f} slot-free {us(43):.2f} converted {us(44):.2f} "
              f"arrived {us(9):.2f} barrier {us(46):.2f} issued {us(47):.2f}")
        print(f"            mma: wait-A {us(51):.2f} got-A {us(52):.2f} committed {us(53):.2f}   "
              f"epilogue: wait-acc {us(48):.2f} got-acc {us(49):.2f} stored {us(50):.2f}")
    if t[54] and t[58]:   # epilogue, first item, second sub-tile (warm code): us since the unit started
        ru = lambda e: (t[e] - t[54]) / 1e3
        print(f"    epilogue unit: acc loaded {ru(55):.2f} stored {ru(56):.2f} next operands requested {ru(57):.2f} statistics {ru(58):.2f}")
    tl[slot] = 0

# fused level-0 kernel (slot 63): one steady-state item (the CTA's 4th), microseconds since that item's start
t = tl[63, 0]
if t[0] and t[24]:
    rel = lambda e: (t[e] - t[24]) / 1e3 if t[e] else float("nan")
    print(f"\nlevel0_fused_kernel: kernel {(t[40] - t[0]) / 1e3:.1f} us; item 3 of CTA 0, us since the item started "
          f"(signals loaded {rel(25):.2f}, a1 stored br0 {rel(26):.2f} br1 {rel(27):.2f})")
    for layer in range(3):
        for br in range(2):
            i = (layer * 2 + br) * 3
            print(f"  layer {layer} br {br}: mma wait {rel(1 + i):.2f} -> go {rel(2 + i):.2f} -> issued {rel(3 + i):.2f} | "
                  f"workers wait {rel(28 + i):.2f} -> acc {rel(29 + i):.2f} -> stored {rel(30 + i):.2f}")
    print(f"  film_out   : mma wait {rel(19):.2f} -> go {rel(20):.2f} -> issued {rel(21):.2f} | workers from {rel(46):.2f} -> stored {rel(48):.2f}")
