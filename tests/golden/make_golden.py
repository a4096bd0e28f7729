"""Generate the golden fixtures in this directory FROM THE REFERENCE ITSELF.

Run in the build container only (needs ``/root/reference``; it does not exist
on the GPU box and nothing at test/bench time reads it):

    python tests/golden/make_golden.py

The reference (`harana`, pure Python/PyTorch) is imported unmodified with empty
stub modules for the optional dependencies it imports at module top but never
uses on this path (h5py, librosa, kaldiio, tkinter -- SURVEY.md 8c).  Weights
and inputs come from ``svcc23_fastsvc_b200.synthetic`` (numpy ``RandomState``
streams), are loaded into the reference modules through ``load_state_dict``,
and the reference's fp32 CPU output is stored.  Fixtures hold
(config, seeds, output) -- inputs/weights are regenerated from the seeds.
"""

import importlib.util
import json
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"


def _load_synthetic():
    # by file path: keeps the repo root (which has its own `harana` drop-in) off sys.path
    spec = importlib.util.spec_from_file_location(
        "fsvc_synthetic", os.path.join(REPO, "svcc23_fastsvc_b200", "synthetic.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _import_reference():
    for name in ("h5py", "librosa", "kaldiio"):
        sys.modules.setdefault(name, types.ModuleType(name))
    tk = types.ModuleType("tkinter")
    tk.W = "w"
    sys.modules.setdefault("tkinter", tk)
    sys.path[:] = [p for p in sys.path if os.path.abspath(p or ".") != REPO]
    sys.path.insert(0, REF)
    import harana.models as models  # noqa: the REFERENCE package
    assert models.__file__.startswith(REF), models.__file__
    return models


def main():
    import torch

    torch.set_num_threads(8)
    syn = _load_synthetic()
    models = _import_reference()
    from harana.models.fastsvc import FastSVCDownsampleNet, FastSVCFiLMNet, FastSVCUpsampleNet

    def t(a):
        return None if a is None else torch.from_numpy(np.ascontiguousarray(a))

    def build_generator(cfg, wseed, weight_norm):
        g = models.FastSVCGenerator(**{k: (list(v) if isinstance(v, (list, tuple)) else v)
                                       for k, v in cfg.items()})
        if not weight_norm:
            g.remove_weight_norm()
        params = syn.make_params(cfg, seed=wseed, weight_norm=weight_norm)
        sd = g.state_dict()
        assert set(sd.keys()) == set(params.keys()), set(sd.keys()) ^ set(params.keys())
        for k in sd:
            assert tuple(sd[k].shape) == params[k].shape, (k, sd[k].shape, params[k].shape)
        g.load_state_dict({k: t(v) for k, v in params.items()})
        return g.eval()

    index = {}

    def run_gen(name, cfg, B, frames, wseed, iseed, with_spk=True, weight_norm=False, subsample=1):
        g = build_generator(cfg, wseed, weight_norm)
        ppg, sine, lft, spk = syn.make_inputs(B, frames, cfg, seed=iseed, with_spk=with_spk)
        with torch.no_grad():
            y = g(t(ppg), t(sine), t(lft), t(spk)).numpy()
            y64 = g.double()(t(ppg).double(), t(sine).double(), t(lft).double(),
                             None if spk is None else t(spk).double()).numpy()
        meta = dict(kind="generator", config=cfg, B=B, frames=frames, wseed=wseed, iseed=iseed,
                    with_spk=with_spk, weight_norm=weight_norm, subsample=subsample,
                    shape=list(y.shape), sum64=float(y.astype(np.float64).sum()),
                    sumsq64=float((y.astype(np.float64) ** 2).sum()),
                    ref32_vs_ref64_maxabs=float(np.abs(y - y64).max()),
                    absmax=float(np.abs(y).max()))
        np.savez_compressed(os.path.join(HERE, name + ".npz"),
                            out=y[..., ::subsample].astype(np.float32),
                            out64=y64[..., ::subsample].astype(np.float32))
        index[name] = meta
        print(name, meta["shape"], "absmax", meta["absmax"], "32v64", meta["ref32_vs_ref64_maxabs"])

    yaml_cfg = dict(syn.YAML_CONFIG)
    run_gen("gen_yaml_b1", yaml_cfg, 1, 100, 0, 1234)                         # config 1
    run_gen("gen_yaml_b1_nospk", yaml_cfg, 1, 100, 0, 1234, with_spk=False)
    run_gen("gen_yaml_b1_wn", yaml_cfg, 1, 100, 1, 1235, weight_norm=True)
    cfg5442 = dict(yaml_cfg, upsampling_scales=[5, 4, 4, 2])                   # README.md:32 variant
    run_gen("gen_5442_b1", cfg5442, 1, 100, 2, 1236)
    run_gen("gen_yaml_b2_f51", yaml_cfg, 2, 51, 0, 1237)                       # config 3/4 shape
    run_gen("gen_yaml_b32", yaml_cfg, 32, 100, 0, 1234, subsample=16)          # config 2
    nospk_cfg = dict(yaml_cfg, use_spk_emb=False)
    run_gen("gen_yaml_nospkmodule_b1", nospk_cfg, 1, 20, 3, 1238, with_spk=False)
    odd_cfg = dict(in_channels=10, mid_channels=[20, 12, 6, 5], upsampling_scales=[3, 2, 2, 3],
                   out_channels=2, spk_emb_size=16, use_spk_emb=True)
    run_gen("gen_odd_b3", odd_cfg, 3, 7, 4, 1239)
    run_gen("gen_yaml_b1_f1", yaml_cfg, 1, 1, 0, 1240)                         # shortest legal input
    run_gen("gen_yaml_b1_f500", yaml_cfg, 1, 500, 0, 1241, subsample=8)        # config 5 shape (5 s)

    # ---- sub-blocks ----
    def load_block(mod, params, prefix):
        sd = {k[len(prefix) + 1:]: t(v) for k, v in params.items() if k.startswith(prefix + ".")}
        mod.load_state_dict(sd)
        return mod.eval()

    def strip_wn(mod):
        for m in mod.modules():
            try:
                torch.nn.utils.remove_weight_norm(m)
            except ValueError:
                pass
        return mod

    params = syn.make_params(yaml_cfg, seed=7)
    rs = np.random.RandomState(99)
    blocks = {}
    # level 0 (1 -> 24, /1) and level 1 (24 -> 48, /5) of the sine branch
    x0 = rs.standard_normal(size=(2, 1, 800)).astype(np.float32)
    d0 = load_block(strip_wn(FastSVCDownsampleNet(1, 24, 1)), params, "downsampling_sine.0")
    d1 = load_block(strip_wn(FastSVCDownsampleNet(24, 48, 5)), params, "downsampling_sine.1")
    f0 = load_block(strip_wn(FastSVCFiLMNet(24)), params, "film_sine.0")
    up = load_block(strip_wn(FastSVCUpsampleNet(48, 24, 5, 512, True)), params, "upsampling_nets.3")
    with torch.no_grad():
        y0 = d0(t(x0))
        y1 = d1(y0)
        sc, sh = f0(y0)
        blocks.update(down0_in=x0, down0_out=y0.numpy(), down1_out=y1.numpy(),
                      film0_scale=sc.numpy(), film0_shift=sh.numpy())
        xin = rs.standard_normal(size=(2, 48, 160)).astype(np.float32)
        g_s = rs.standard_normal(size=(2, 24, 800)).astype(np.float32)
        b_s = rs.standard_normal(size=(2, 24, 800)).astype(np.float32)
        g_l = rs.standard_normal(size=(2, 24, 800)).astype(np.float32)
        b_l = rs.standard_normal(size=(2, 24, 800)).astype(np.float32)
        spk = rs.standard_normal(size=(2, 512)).astype(np.float32)
        yu = up(t(xin), (t(g_s), t(b_s)), (t(g_l), t(b_l)), t(spk))
        yu_nospk = up(t(xin), (t(g_s), t(b_s)), (t(g_l), t(b_l)), None)
        blocks.update(up_in=xin, up_gs=g_s, up_bs=b_s, up_gl=g_l, up_bl=b_l, up_spk=spk,
                      up_out=yu.numpy(), up_out_nospk=yu_nospk.numpy())
    np.savez_compressed(os.path.join(HERE, "blocks.npz"), **blocks)
    index["blocks"] = dict(kind="blocks", wseed=7, config=yaml_cfg)
    print("blocks", {k: v.shape for k, v in blocks.items()})

    # ---- layer semantics (a7, a8): Stretch2d / Squeeze2d incl. the non-divisible case ----
    from harana.layers import Squeeze2d, Stretch2d
    lay = {}
    v = np.arange(23, dtype=np.float32).reshape(1, 1, 23)
    lay["squeeze_23_5"] = Squeeze2d(5)(t(v)).numpy()
    lay["squeeze_40_4"] = Squeeze2d(4)(t(np.arange(40, dtype=np.float32).reshape(1, 1, 40))).numpy()
    lay["stretch_7_5"] = Stretch2d(5, 1)(t(np.arange(7, dtype=np.float32).reshape(1, 1, 1, 7))).numpy()
    np.savez_compressed(os.path.join(HERE, "layers.npz"), **lay)
    index["layers"] = dict(kind="layers")

    # state_dict key sets (boundary contract)
    g = models.FastSVCGenerator()
    keys_wn = {k: list(v.shape) for k, v in g.state_dict().items()}
    g.remove_weight_norm()
    keys_plain = {k: list(v.shape) for k, v in g.state_dict().items()}
    with open(os.path.join(HERE, "state_dict_keys.json"), "w") as f:
        json.dump(dict(weight_norm=keys_wn, plain=keys_plain, repr=repr(g)), f, indent=0)

    with open(os.path.join(HERE, "index.json"), "w") as f:
        json.dump(index, f, indent=1)
    print("wrote", len(index), "fixtures to", HERE)


if __name__ == "__main__":
    main()
