"""The oracle (oracle/fastsvc_numpy.py, oracle/fastsvc_torch.py) against the
golden vectors generated from the reference itself (tests/golden/make_golden.py).
CPU only."""
import numpy as np
import pytest
import torch

from conftest import BIG_CASES, GENERATOR_CASES, case_inputs, load_golden
from oracle import fastsvc_numpy as onp
from oracle import fastsvc_torch as otorch
from svcc23_fastsvc_b200 import synthetic as syn

TOL = 1e-4  # oracle vs reference fp32 output, max-abs (reference fp32-vs-fp64 floor is ~1.5e-5)


@pytest.mark.parametrize("name", [c for c in GENERATOR_CASES if c not in BIG_CASES])
def test_numpy_oracle_matches_reference(name, golden_index):
    meta = golden_index[name]
    params, ppg, sine, lft, spk = case_inputs(meta)
    y = onp.generator_forward(params, ppg, sine, lft, spk, meta["config"]["upsampling_scales"])
    gold = load_golden(name)["out"]
    assert list(y.shape) == meta["shape"]
    assert np.abs(y[..., ::meta["subsample"]] - gold).max() <= TOL


@pytest.mark.parametrize("name", GENERATOR_CASES)
@pytest.mark.parametrize("recompute", [True, False])
def test_torch_port_matches_reference(name, recompute, golden_index):
    if name in BIG_CASES and not recompute:
        pytest.skip("one pass over the big case is enough")
    meta = golden_index[name]
    params, ppg, sine, lft, spk = case_inputs(meta)
    tp = {k: torch.from_numpy(v) for k, v in params.items()}
    with torch.no_grad():
        y = otorch.generator_forward(tp, torch.from_numpy(ppg), torch.from_numpy(sine), torch.from_numpy(lft),
                                     None if spk is None else torch.from_numpy(spk),
                                     meta["config"]["upsampling_scales"], recompute=recompute).numpy()
    gold = load_golden(name)["out"]
    assert list(y.shape) == meta["shape"]
    assert np.abs(y[..., ::meta["subsample"]] - gold).max() <= TOL
    if meta["subsample"] == 1:
        assert abs(float(y.astype(np.float64).sum()) - meta["sum64"]) <= 1e-2 * max(1.0, abs(meta["sum64"]))


def test_numpy_oracle_fp64_matches_reference_fp64(golden_index):
    meta = golden_index["gen_yaml_b1"]
    params, ppg, sine, lft, spk = case_inputs(meta)
    p64 = {k: v.astype(np.float64) for k, v in params.items()}
    y = onp.generator_forward(p64, ppg.astype(np.float64), sine.astype(np.float64), lft.astype(np.float64),
                              spk.astype(np.float64))
    assert np.abs(y - load_golden("gen_yaml_b1")["out64"]).max() <= 2e-6  # out64 is stored rounded to fp32


def test_blocks_match_reference():
    g = load_golden("blocks")
    params = syn.make_params(syn.YAML_CONFIG, seed=7)
    y0 = onp.downsample_net(params, "downsampling_sine.0", g["down0_in"], 1)
    assert np.abs(y0 - g["down0_out"]).max() <= 1e-5
    y1 = onp.downsample_net(params, "downsampling_sine.1", y0, 5)
    assert np.abs(y1 - g["down1_out"]).max() <= 1e-5
    sc, sh = onp.film_net(params, "film_sine.0", y0)
    assert np.abs(sc - g["film0_scale"]).max() <= 1e-5
    assert np.abs(sh - g["film0_shift"]).max() <= 1e-5
    for spk, key in ((g["up_spk"], "up_out"), (None, "up_out_nospk")):
        yu = onp.upsample_net(params, "upsampling_nets.3", g["up_in"], (g["up_gs"], g["up_bs"]),
                              (g["up_gl"], g["up_bl"]), 5, spk)
        assert np.abs(yu - g[key]).max() <= 1e-4


def test_layer_semantics_match_reference():
    g = load_golden("layers")
    v = np.arange(23, dtype=np.float32).reshape(1, 1, 23)
    assert np.array_equal(onp.squeeze(v, 5), g["squeeze_23_5"])          # non-divisible: idx 0,5,11,17
    assert np.array_equal(onp.squeeze(np.arange(40, dtype=np.float32).reshape(1, 1, 40), 4), g["squeeze_40_4"])
    assert np.array_equal(onp.stretch(np.arange(7, dtype=np.float32).reshape(1, 1, 1, 7), 5), g["stretch_7_5"])


def test_weight_norm_folding():
    p = syn.make_params(syn.YAML_CONFIG, seed=3, weight_norm=True)
    w = onp.effective_weight(p, "film_lft.2.conv")
    v, gg = p["film_lft.2.conv.weight_v"], p["film_lft.2.conv.weight_g"]
    n = np.sqrt((v.astype(np.float64) ** 2).sum(axis=(1, 2), keepdims=True))
    assert np.allclose(w, gg * v / n, rtol=1e-5, atol=1e-7)
