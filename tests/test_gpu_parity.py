"""Parity of the CUDA path (through harana.models -> ctypes -> libfsvc.so C ABI) against the golden
vectors generated from the reference and against the CPU oracle.  Needs a B200: run with -m gpu."""
import numpy as np
import pytest
import torch

from conftest import GENERATOR_CASES, case_inputs, load_golden

pytestmark = pytest.mark.gpu

# BASELINE.json north_star: waveform parity <= 1e-3 max-abs vs the reference fp32 generator.
TOL = 1e-3
MODES = ["fp32", "tc_bf16x3", "auto"]


def _cuda():
    assert torch.cuda.is_available(), "gpu tests need a CUDA device"
    return torch.device("cuda:0")


def _build_generator(meta, precision):
    import harana.models as M
    params, ppg, sine, lft, spk = case_inputs(meta)
    g = M.FastSVCGenerator(**{k: (list(v) if isinstance(v, list) else v) for k, v in meta["config"].items()})
    if not meta["weight_norm"]:
        g.remove_weight_norm()
    g.load_state_dict({k: torch.from_numpy(v) for k, v in params.items()})
    g = g.eval().to(_cuda())
    g.precision = precision
    return g, params, ppg, sine, lft, spk


def _t(a):
    return None if a is None else torch.from_numpy(a).to(_cuda())


@pytest.mark.parametrize("precision", MODES)
@pytest.mark.parametrize("name", GENERATOR_CASES)
def test_generator_matches_reference_golden(name, precision, golden_index):
    meta = golden_index[name]
    g, params, ppg, sine, lft, spk = _build_generator(meta, precision)
    with torch.no_grad():
        y = g(_t(ppg), _t(sine), _t(lft), _t(spk))
    assert y.is_cuda and list(y.shape) == meta["shape"] and y.dtype == torch.float32
    y = y.cpu().numpy()
    assert np.isfinite(y).all()
    gold = load_golden(name)["out"]
    err = np.abs(y[..., ::meta["subsample"]] - gold).max()
    assert err <= TOL, f"{name}/{precision}: max-abs {err:.3e} (absmax {meta['absmax']:.2f})"
    assert g.last_launch_count() > 0


@pytest.mark.parametrize("precision", MODES)
def test_generator_matches_oracle_on_fresh_inputs(precision, golden_index):
    """Seeds the golden set does not contain; checked against the numpy oracle run here."""
    from oracle import fastsvc_numpy as onp
    from svcc23_fastsvc_b200 import synthetic as syn
    meta = dict(golden_index["gen_yaml_b1"], wseed=11, iseed=77, B=3, frames=13)
    g, params, ppg, sine, lft, spk = _build_generator(meta, precision)
    ref = onp.generator_forward(params, ppg, sine, lft, spk)
    with torch.no_grad():
        y = g(_t(ppg), _t(sine), _t(lft), _t(spk)).cpu().numpy()
    assert np.abs(y - ref).max() <= TOL


@pytest.mark.parametrize("precision", ["fp32", "tc_bf16x3"])
def test_blocks_match_reference_golden(precision):
    from harana.models.fastsvc import FastSVCDownsampleNet, FastSVCFiLMNet, FastSVCUpsampleNet
    from svcc23_fastsvc_b200 import synthetic as syn
    gold = load_golden("blocks")
    params = syn.make_params(syn.YAML_CONFIG, seed=7)

    def load(mod, prefix):
        mod.load_state_dict({k[len(prefix) + 1:]: torch.from_numpy(v) for k, v in params.items()
                             if k.startswith(prefix + ".")})
        mod.precision = precision
        return mod.eval().to(_cuda())

    d0 = load(FastSVCDownsampleNet(1, 24, 1), "downsampling_sine.0")
    d1 = load(FastSVCDownsampleNet(24, 48, 5), "downsampling_sine.1")
    f0 = load(FastSVCFiLMNet(24), "film_sine.0")
    up = load(FastSVCUpsampleNet(48, 24, 5, 512, True), "upsampling_nets.3")
    with torch.no_grad():
        y0 = d0(_t(gold["down0_in"]))
        y1 = d1(y0)
        sc, sh = f0(y0)
        yu = up(_t(gold["up_in"]), (_t(gold["up_gs"]), _t(gold["up_bs"])), (_t(gold["up_gl"]), _t(gold["up_bl"])),
                _t(gold["up_spk"]))
        yn = up(_t(gold["up_in"]), (_t(gold["up_gs"]), _t(gold["up_bs"])), (_t(gold["up_gl"]), _t(gold["up_bl"])),
                None)
    for got, key in ((y0, "down0_out"), (y1, "down1_out"), (sc, "film0_scale"), (sh, "film0_shift"),
                     (yu, "up_out"), (yn, "up_out_nospk")):
        err = np.abs(got.cpu().numpy() - gold[key]).max()
        assert err <= (1e-4 if precision == "fp32" else 1e-3), (key, err)
    with torch.no_grad(), pytest.raises(ValueError):
        d1(torch.zeros(1, 24, 23, device=_cuda()))  # T not divisible by the scale
    with pytest.raises(RuntimeError, match="inference-only"):   # grad-enabled call on a block: refused, not dropped
        f0(y0)


@pytest.mark.parametrize("precision", MODES)
def test_full_size_properties(precision, golden_index):
    """BASELINE config 2 (B=32, 1-s clips): size-independent properties."""
    meta = golden_index["gen_yaml_b32"]
    g, params, ppg, sine, lft, spk = _build_generator(meta, precision)
    x, s, l, e = _t(ppg), _t(sine), _t(lft), _t(spk)
    with torch.no_grad():
        y = g(x, s, l, e)
        y_again = g(x, s, l, e)
        perm = torch.randperm(32, device=x.device, generator=torch.Generator(device=x.device).manual_seed(0))
        y_perm = g(x[perm].contiguous(), s[perm].contiguous(), l[perm].contiguous(), e[perm].contiguous())
        y_one = g(x[5:6].contiguous(), s[5:6].contiguous(), l[5:6].contiguous(), e[5:6].contiguous())
    assert torch.equal(y, y_again), "forward must be run-to-run deterministic"
    # utterances are independent (InstanceNorm is per (b, c)): batch order / batch size change nothing
    assert torch.equal(y[perm], y_perm)
    assert torch.equal(y[5:6], y_one)
    # checksum against the reference's own output of the whole batch
    ysum = float(y.double().sum())
    assert abs(ysum - meta["sum64"]) <= 1e-3 * 32 * 16000 * 1e-2 + 1e-4 * abs(meta["sum64"])


def test_operand_planes_change_no_bit(golden_index, monkeypatch):
    """Levels whose CTAs see few items hand d1 -> d2 -> d4 and film_conv -> film_out over as bf16 hi|lo operand planes
    (conv_tc3 MODE 6: the producer's epilogue applies the consumer's LeakyReLU and split).  Same arithmetic in another
    place: the waveform must not change by a bit against the transform path (FSVC_NO_PLANES=1)."""
    for case in ("gen_yaml_b2_f51", "gen_yaml_b32"):
        meta = golden_index[case]
        g, params, ppg, sine, lft, spk = _build_generator(meta, "tc_bf16x3")
        args = [_t(ppg), _t(sine), _t(lft), _t(spk)]
        with torch.no_grad():
            monkeypatch.delenv("FSVC_NO_PLANES", raising=False)
            y_planes = g(*args)
            monkeypatch.setenv("FSVC_NO_PLANES", "1")
            y_transform = g(*args)
            monkeypatch.delenv("FSVC_NO_PLANES", raising=False)
        assert torch.equal(y_planes, y_transform), case


def test_forward_host_matches_device_forward(golden_index):
    meta = golden_index["gen_yaml_b2_f51"]
    g, params, ppg, sine, lft, spk = _build_generator(meta, "auto")
    with torch.no_grad():
        y_dev = g(_t(ppg), _t(sine), _t(lft), _t(spk)).cpu()
        pin = [torch.from_numpy(a).pin_memory() for a in (ppg, sine, lft, spk)]
        y_host = g.forward_host(*pin)
        torch.cuda.synchronize()
    assert torch.equal(y_dev, y_host)


def test_weight_updates_are_picked_up(golden_index):
    meta = golden_index["gen_yaml_b1_f1"]
    g, params, ppg, sine, lft, spk = _build_generator(meta, "auto")
    with torch.no_grad():
        y0 = g(_t(ppg), _t(sine), _t(lft), _t(spk)).clone()
        g.conv_last.bias.add_(1.0)          # in-place update, like an optimizer step
        y1 = g(_t(ppg), _t(sine), _t(lft), _t(spk))
    assert torch.allclose(y1, y0 + 1.0, atol=1e-5)


def test_argument_errors(golden_index):
    meta = golden_index["gen_yaml_b1_f1"]
    g, params, ppg, sine, lft, spk = _build_generator(meta, "auto")
    with pytest.raises(ValueError):
        g(_t(ppg), _t(sine)[..., :-1].contiguous(), _t(lft), _t(spk))
    with pytest.raises(ValueError):
        g(_t(ppg)[:, :100].contiguous(), _t(sine), _t(lft), _t(spk))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        g(torch.from_numpy(ppg), torch.from_numpy(sine), torch.from_numpy(lft), None)


def test_inference_signature(golden_index):
    """FastSVCGenerator.inference (fastsvc.py:364-383) as decode_fastsvc.py:187-189 calls it."""
    meta = golden_index["gen_yaml_b1"]
    g, params, ppg, sine, lft, spk = _build_generator(meta, "auto")
    sig = _t(sine)
    with torch.no_grad():
        y_fwd = g(_t(ppg), sig, _t(lft), _t(spk))
        y_inf = g.inference(_t(ppg)[0].t().contiguous(), torch.zeros(100, 1, device=_cuda()), _t(lft)[0].t().contiguous(),
                            lambda f0: sig, torch.nn.ReplicationPad1d(0), _t(spk))
    assert y_inf.shape == (16000, 1)
    assert torch.equal(y_inf[:, 0], y_fwd[0, 0])


def test_training_step_gradients_match_torch_graph():
    """Grad-enabled forward (train_fastsvc.py:168,199-206): values from CUDA kernels, gradients finite and
    equal to those of the oracle graph."""
    import harana.models as M
    from oracle import fastsvc_torch as otorch
    from svcc23_fastsvc_b200 import synthetic as syn
    cfg = dict(in_channels=16, mid_channels=[16, 8], upsampling_scales=[2, 3], out_channels=1, spk_emb_size=8,
               use_spk_emb=True)
    torch.backends.cudnn.allow_tf32 = False     # the comparison graph must be true fp32
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.manual_seed(0)                        # weights come from the module's own random init
    g = M.FastSVCGenerator(**cfg).to(_cuda())
    ppg, sine, lft, spk = syn.make_inputs(2, 6, cfg, seed=3)
    y = g(_t(ppg), _t(sine), _t(lft), _t(spk))
    assert y.requires_grad
    y.square().mean().backward()
    grads = {k: p.grad.clone() for k, p in g.named_parameters()}
    assert all(torch.isfinite(v).all() for v in grads.values())
    sd = {k: v.detach().clone().requires_grad_(True) for k, v in g.state_dict().items()}
    yo = otorch.generator_forward(sd, _t(ppg), _t(sine), _t(lft), _t(spk), cfg["upsampling_scales"], recompute=False)
    assert torch.allclose(y, yo, atol=TOL)      # forward values: the parity bar (tensor-core mode)
    yo.square().mean().backward()
    for k, v in grads.items():
        assert torch.allclose(v, sd[k].grad, atol=1e-4, rtol=1e-3), k
