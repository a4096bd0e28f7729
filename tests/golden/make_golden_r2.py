"""Round-2 golden fixtures, generated FROM THE REFERENCE ITSELF like make_golden.py (build container only):

    python tests/golden/make_golden_r2.py

Wider tensor-core-mode coverage asked for by the round-1 review: other multiple-of-8 channel sets, a level-0 width
that does not take the fused level kernel, batch > 1 at 500 frames (statistics merged by the separate kernel) and at
33 frames (ragged last tile), and the (1, S) target-speaker broadcast of decode_fastsvc.py:156-158.
Entries are appended to index.json; inputs / weights are regenerated from the seeds (conftest.case_inputs).
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as mg  # noqa: E402


def main():
    import torch
    torch.set_num_threads(8)
    syn = mg._load_synthetic()
    models = mg._import_reference()
    with open(os.path.join(HERE, "index.json")) as f:
        index = json.load(f)

    def t(a):
        return None if a is None else torch.from_numpy(np.ascontiguousarray(a))

    def run_gen(name, cfg, B, frames, wseed, iseed, spk_rows=None, subsample=1):
        g = models.FastSVCGenerator(**{k: (list(v) if isinstance(v, (list, tuple)) else v) for k, v in cfg.items()})
        g.remove_weight_norm()
        params = syn.make_params(cfg, seed=wseed)
        g.load_state_dict({k: t(v) for k, v in params.items()})
        g = g.eval()
        ppg, sine, lft, spk = syn.make_inputs(B, frames, cfg, seed=iseed)
        if spk_rows is not None:
            spk = spk[:spk_rows]
        with torch.no_grad():
            y = g(t(ppg), t(sine), t(lft), t(spk)).numpy()
            y64 = g.double()(t(ppg).double(), t(sine).double(), t(lft).double(), t(spk).double()).numpy()
        meta = dict(kind="generator", config=cfg, B=B, frames=frames, wseed=wseed, iseed=iseed, with_spk=True,
                    weight_norm=False, subsample=subsample, shape=list(y.shape),
                    sum64=float(y.astype(np.float64).sum()), sumsq64=float((y.astype(np.float64) ** 2).sum()),
                    ref32_vs_ref64_maxabs=float(np.abs(y - y64).max()), absmax=float(np.abs(y).max()))
        if spk_rows is not None:
            meta["spk_rows"] = spk_rows
        np.savez_compressed(os.path.join(HERE, name + ".npz"), out=y[..., ::subsample].astype(np.float32),
                            out64=y64[..., ::subsample].astype(np.float32))
        index[name] = meta
        print(name, meta["shape"], "absmax", meta["absmax"], "32v64", meta["ref32_vs_ref64_maxabs"])

    yaml_cfg = dict(syn.YAML_CONFIG)
    run_gen("gen_c64_b2", dict(yaml_cfg, in_channels=80, mid_channels=[64, 32, 16, 8], spk_emb_size=64), 2, 20, 5, 1301)
    run_gen("gen_c256_b1", dict(yaml_cfg, mid_channels=[256, 128, 64, 32]), 1, 12, 6, 1302)
    run_gen("gen_c48last_b2", dict(yaml_cfg, mid_channels=[192, 96, 48, 48]), 2, 10, 7, 1303)
    run_gen("gen_c40last_b1", dict(yaml_cfg, mid_channels=[96, 48, 40, 40], spk_emb_size=32), 1, 9, 8, 1304)
    run_gen("gen_yaml_b3_f500", yaml_cfg, 3, 500, 0, 1305, subsample=8)
    run_gen("gen_yaml_b2_f33", yaml_cfg, 2, 33, 0, 1306)
    run_gen("gen_yaml_b4_spk1", yaml_cfg, 4, 40, 0, 1307, spk_rows=1)
    run_gen("gen_3stage_b2", dict(yaml_cfg, mid_channels=[64, 32, 16], upsampling_scales=[4, 8, 5]), 2, 15, 9, 1308)

    with open(os.path.join(HERE, "index.json"), "w") as f:
        json.dump(index, f, indent=1)


if __name__ == "__main__":
    main()
