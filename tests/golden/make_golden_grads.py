"""Gradient fingerprints of the REFERENCE generator (training path, train_fastsvc.py:168,199-206), for SURVEY 8f N1.

    python tests/golden/make_golden_grads.py        # build container only (needs /root/reference)

For each case: loss = sum(y * w) with a seeded w; the reference's autograd gradients of every parameter are reduced
to (L2 norm, sum, 8 sampled entries) -- 2.7 M gradient values do not belong in the repo.  tests/test_grads.py checks
the oracle's torch port against them on CPU and the CUDA generator's grad-enabled forward on the GPU."""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import _import_reference, _load_synthetic  # noqa: E402


def fingerprint(g):
    g = np.asarray(g, dtype=np.float64).reshape(-1)
    idx = np.linspace(0, g.size - 1, num=min(8, g.size)).astype(np.int64)
    return dict(norm=float(np.sqrt((g * g).sum())), sum=float(g.sum()), idx=idx.tolist(), val=g[idx].tolist())


def main():
    import torch
    torch.set_num_threads(8)
    syn = _load_synthetic()
    models = _import_reference()
    cases = {
        "grad_yaml_b2_f8": dict(cfg=dict(syn.YAML_CONFIG), B=2, frames=8, wseed=0, iseed=77, with_spk=True),
        "grad_yaml_b1_f5_nospk": dict(cfg=dict(syn.YAML_CONFIG), B=1, frames=5, wseed=0, iseed=78, with_spk=False),
        # round 2: weight norm applied (what the reference trains with, fastsvc.py:303), the config-3/4 segment length
        # (51 frames = 8160 samples: several time tiles / weight-gradient splits), odd channel counts + 2 outputs
        "grad_yaml_b2_f6_wn": dict(cfg=dict(syn.YAML_CONFIG), B=2, frames=6, wseed=3, iseed=79, with_spk=True,
                                   weight_norm=True),
        "grad_yaml_b2_f51": dict(cfg=dict(syn.YAML_CONFIG), B=2, frames=51, wseed=0, iseed=80, with_spk=True),
        "grad_odd_b3_f7": dict(cfg=dict(in_channels=10, mid_channels=[20, 12, 6, 5], upsampling_scales=[3, 2, 2, 3],
                                        out_channels=2, spk_emb_size=16, use_spk_emb=True),
                               B=3, frames=7, wseed=4, iseed=81, with_spk=True),
    }
    out = {}
    for name, cs in cases.items():
        cfg = cs["cfg"]
        g = models.FastSVCGenerator(**{k: (list(v) if isinstance(v, (list, tuple)) else v) for k, v in cfg.items()})
        wn = cs.get("weight_norm", False)
        if not wn:
            g.remove_weight_norm()
        params = syn.make_params(cfg, seed=cs["wseed"], weight_norm=wn)
        g.load_state_dict({k: torch.from_numpy(v) for k, v in params.items()})
        g.train()
        ppg, sine, lft, spk = syn.make_inputs(cs["B"], cs["frames"], cfg, seed=cs["iseed"], with_spk=cs["with_spk"])
        T = cs["frames"] * syn.hop_size(cfg["upsampling_scales"])
        w = np.random.RandomState(cs["iseed"] + 1000).standard_normal(size=(cs["B"], cfg["out_channels"], T)).astype(np.float32)
        t = lambda a: None if a is None else torch.from_numpy(a)
        y = g(t(ppg), t(sine), t(lft), t(spk))
        (y * torch.from_numpy(w)).sum().backward()
        fp = {k: fingerprint(p.grad.numpy()) for k, p in g.named_parameters() if p.grad is not None}
        out[name] = dict(B=cs["B"], frames=cs["frames"], wseed=cs["wseed"], iseed=cs["iseed"], with_spk=cs["with_spk"],
                         weight_norm=wn, config=cfg, loss=float((y.detach() * torch.from_numpy(w)).sum()), grads=fp)
        print(name, len(fp), "parameter gradients; loss", out[name]["loss"])
    with open(os.path.join(HERE, "grads.json"), "w") as f:
        json.dump(out, f)


if __name__ == "__main__":
    main()
