"""Training path on the GPU: the reference's _train_step order (train_fastsvc.py:157-235)
driven by GanTrainer with the native generator, a host-PyTorch critic and the MR-STFT / LSGAN losses; plus the
autograd-contract checks of the native backward."""
import os
import sys

import numpy as np
import pytest
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "tools"))

pytestmark = pytest.mark.gpu


def _gen(dev, weight_norm=True, seed=0):
    import harana.models as M
    from svcc23_fastsvc_b200 import synthetic as syn
    cfg = dict(syn.YAML_CONFIG)
    params = syn.make_params(cfg, seed=seed, weight_norm=weight_norm)
    g = M.FastSVCGenerator(**{k: (list(v) if isinstance(v, list) else v) for k, v in cfg.items()})
    if not weight_norm:
        g.remove_weight_norm()
    g.load_state_dict({k: torch.from_numpy(v) for k, v in params.items()})
    return g.train().to(dev), params


def test_two_training_steps_like_the_reference_trainer():
    import gan_host
    from svcc23_fastsvc_b200 import synthetic as syn
    from svcc23_fastsvc_b200.training import GanTrainer
    dev = torch.device("cuda:0")
    g, _ = _gen(dev)
    torch.manual_seed(0)
    D = gan_host.MultiScaleMultiPeriodCritic().to(dev)
    stft = gan_host.MultiResolutionSTFTLoss(**gan_host.STFT_PARAMS).to(dev)
    opt_g = torch.optim.RAdam(g.parameters(), lr=1e-3, eps=1e-6)
    opt_d = torch.optim.RAdam(D.parameters(), lr=1e-3, eps=1e-6)
    tr = GanTrainer(g, D, stft, gan_host.generator_adversarial_loss, gan_host.discriminator_adversarial_loss, opt_g,
                    opt_d, lambda_adv=2.5, generator_grad_norm=10.0, discriminator_grad_norm=1.0)
    B, frames = 2, 51                                              # the collater's 8160-sample segments
    ins = [torch.from_numpy(a).to(dev) for a in syn.make_inputs(B, frames, syn.YAML_CONFIG, seed=5)]
    y = 0.1 * torch.randn(B, 1, frames * 160, device=dev)
    before = [p.detach().clone() for p in g.parameters()]
    d_before = [p.detach().clone() for p in D.parameters()]
    for _ in range(2):
        logs = tr.step(tuple(ins), y, adversarial=True)
    tr.finish_discriminator_step()
    torch.cuda.synchronize()
    assert all(torch.isfinite(v).all() for v in logs.values()) and float(logs["generator_loss"]) > 0
    assert float(logs["generator_grad_norm"]) > 0
    changed = sum(int(not torch.equal(a, b)) for a, b in zip(before, g.parameters()))
    # every generator parameter got a gradient; all but the structurally gradient-free ones moved (a bias in front of
    # an InstanceNorm -- the 8 FiLM conv_shift biases -- has a rounding-noise gradient, and RAdam's first, unrectified
    # steps scale with the gradient)
    assert changed >= len(before) - 12, changed
    assert any(not torch.equal(a, b) for a, b in zip(d_before, D.parameters()))
    for p in g.parameters():                                       # gradients live in the flat bucket
        assert p.grad is not None and torch.isfinite(p.grad).all()
    tr.gb.check_views()
    # eval-mode inference after training steps sees the updated weights (weights are re-pushed on version change)
    g.eval()
    with torch.no_grad():
        out = g(*ins)
    assert torch.isfinite(out).all()


def test_native_backward_matches_torch_autograd_of_the_oracle_port():
    """Same weights, same inputs: gradients of the native backward vs torch autograd through the oracle's op sequence
    on the GPU in fp32 (no TF32), every parameter, weight norm applied."""
    from oracle import fastsvc_torch as otorch
    from svcc23_fastsvc_b200 import synthetic as syn
    dev = torch.device("cuda:0")
    g, params = _gen(dev, weight_norm=True, seed=3)
    ins = [torch.from_numpy(a).to(dev) for a in syn.make_inputs(3, 20, syn.YAML_CONFIG, seed=9)]
    w = torch.randn(3, 1, 20 * 160, device=dev)
    (g(*ins) * w).sum().backward()
    ours = {k: p.grad.detach().clone() for k, p in g.named_parameters()}
    tp = {k: torch.from_numpy(v).to(dev).requires_grad_(True) for k, v in params.items()}
    with torch.backends.cudnn.flags(enabled=True, allow_tf32=False):
        torch.backends.cuda.matmul.allow_tf32 = False
        (otorch.generator_forward(tp, *ins, recompute=True) * w).sum().backward()
    assert set(ours) == set(tp)
    for k, t in tp.items():
        ref, got = t.grad, ours[k]
        scale = float(ref.norm()) / np.sqrt(ref.numel()) + 1e-6
        # absolute floor: gradients that are structurally zero (a bias feeding InstanceNorm) are rounding noise of a
        # few 1e-4 on both sides (same floor as tests/test_grads.py)
        assert float((got - ref).abs().max()) <= 5e-3 * scale + 1e-3, (k, float((got - ref).abs().max()), scale)


def test_backward_contract_errors():
    from svcc23_fastsvc_b200 import synthetic as syn
    dev = torch.device("cuda:0")
    g, _ = _gen(dev, weight_norm=False)
    ins = [torch.from_numpy(a).to(dev) for a in syn.make_inputs(1, 4, syn.YAML_CONFIG, seed=1)]
    x = ins[0].clone().requires_grad_(True)
    with pytest.raises(NotImplementedError, match="input"):
        g(x, *ins[1:])
    y1 = g(*ins)
    y2 = g(*ins)                                   # a second grad-enabled forward pushes weights again
    with pytest.raises(RuntimeError, match="changed between this forward and its backward"):
        y1.sum().backward()
    y2.sum().backward()                            # the most recent forward can still run its backward
    assert all(p.grad is not None for p in g.parameters())


def test_forward_host_expands_a_single_speaker_row():
    """(1, S) target speaker with B > 1 through fsvc_forward_host (decode_fastsvc.py:156-158)."""
    from svcc23_fastsvc_b200 import synthetic as syn
    dev = torch.device("cuda:0")
    g, _ = _gen(dev, weight_norm=False)
    g.eval()
    ppg, sine, lft, spk = [torch.from_numpy(a) for a in syn.make_inputs(3, 12, syn.YAML_CONFIG, seed=2)]
    with torch.no_grad():
        want = g(ppg.to(dev), sine.to(dev), lft.to(dev), spk[:1].expand(3, -1).contiguous().to(dev))
        got = g.forward_host(ppg, sine, lft, spk[:1])
    torch.cuda.synchronize()
    assert torch.equal(got, want.cpu())
