"""Training path: parameter gradients against fingerprints generated from the REFERENCE's autograd
(tests/golden/grads.json, tests/golden/make_golden_grads.py).  CPU: the oracle's torch port.  GPU: the CUDA generator's
grad-enabled forward (values from libfsvc.so; the backward is the interim PyTorch-op graph, autograd.py)."""
import json
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN

REL_NORM = 5e-4      # relative error of each gradient's L2 norm
REL_ENTRY = 2e-3     # sampled entries, relative to the gradient's RMS magnitude (+ REL_NORM * |entry|)
# Absolute floors: gradients that are structurally zero (a bias feeding InstanceNorm, which removes per-channel
# constants) are rounding noise of a few 1e-4 in the reference itself; real gradients here have norms of 1e-2 .. 1e+1.
ABS_NORM = 1e-3
ABS_ENTRY = 2e-4

with open(os.path.join(GOLDEN, "grads.json")) as f:
    CASES = json.load(f)


def _check(name, grads, want):
    assert set(grads) == set(want), set(grads) ^ set(want)
    for k, fp in want.items():
        g = grads[k].astype(np.float64).reshape(-1)
        norm = np.sqrt((g * g).sum())
        assert abs(norm - fp["norm"]) <= REL_NORM * fp["norm"] + ABS_NORM, (name, k, norm, fp["norm"])
        rms = fp["norm"] / np.sqrt(g.size)
        for i, v in zip(fp["idx"], fp["val"]):
            assert abs(g[i] - v) <= REL_ENTRY * rms + REL_NORM * abs(v) + ABS_ENTRY, (name, k, i, g[i], v)


def _inputs(cs):
    from svcc23_fastsvc_b200 import synthetic as syn
    params = syn.make_params(cs["config"], seed=cs["wseed"], weight_norm=cs.get("weight_norm", False))
    ins = syn.make_inputs(cs["B"], cs["frames"], cs["config"], seed=cs["iseed"], with_spk=cs["with_spk"])
    T = cs["frames"] * syn.hop_size(cs["config"]["upsampling_scales"])
    w = np.random.RandomState(cs["iseed"] + 1000).standard_normal(
        size=(cs["B"], cs["config"]["out_channels"], T)).astype(np.float32)
    return params, ins, w


@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_gradients_match_reference(name):
    from oracle import fastsvc_torch as otorch
    cs = CASES[name]
    params, ins, w = _inputs(cs)
    tp = {k: torch.from_numpy(v).clone().requires_grad_(True) for k, v in params.items()}
    args = [None if a is None else torch.from_numpy(a) for a in ins]
    y = otorch.generator_forward(tp, *args, cs["config"]["upsampling_scales"], recompute=True)
    loss = (y * torch.from_numpy(w)).sum()
    assert abs(float(loss.detach()) - cs["loss"]) <= 1e-4 * max(1.0, abs(cs["loss"]))
    loss.backward()
    grads = {k: v.grad.numpy() for k, v in tp.items() if v.grad is not None and k in cs["grads"]}
    _check(name, grads, cs["grads"])


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(CASES))
def test_cuda_generator_gradients_match_reference(name):
    import harana.models as M
    cs = CASES[name]
    params, ins, w = _inputs(cs)
    dev = torch.device("cuda:0")
    g = M.FastSVCGenerator(**{k: (list(v) if isinstance(v, list) else v) for k, v in cs["config"].items()})
    if not cs.get("weight_norm", False):
        g.remove_weight_norm()
    g.load_state_dict({k: torch.from_numpy(v) for k, v in params.items()})
    g = g.train().to(dev)
    args = [None if a is None else torch.from_numpy(a).to(dev) for a in ins]
    y = g(*args)
    assert y.requires_grad
    loss = (y * torch.from_numpy(w).to(dev)).sum()
    assert abs(float(loss.detach()) - cs["loss"]) <= 2e-3 * w.size * 0.05 + 1e-3   # forward within 1e-3 per sample
    loss.backward()
    grads = {k: p.grad.cpu().numpy() for k, p in g.named_parameters() if p.grad is not None}
    _check(name, grads, cs["grads"])
