"""Drop-in boundary: class lookup, constructor, state_dict layout, weight norm. CPU only."""
import copy
import json
import os

import pytest
import torch

from conftest import GOLDEN


@pytest.fixture(scope="module")
def ref_keys():
    with open(os.path.join(GOLDEN, "state_dict_keys.json")) as f:
        return json.load(f)


def test_lookup_by_name_like_the_reference_scripts():
    import harana.models
    from harana.layers import Conv1d1x1, Conv1d1x3, Conv2d1x3, Squeeze2d, Stretch2d  # noqa: F401
    from harana.models.fastsvc import FastSVCFiLMNet  # noqa: F401  (tacotron2.py:22)
    cls = getattr(harana.models, "FastSVCGenerator")     # train_fastsvc.py:700-713
    params = dict(in_channels=144, out_channels=1, mid_channels=[192, 96, 48, 24],
                  upsampling_scales=[2, 4, 4, 5], spk_emb_size=512, use_spk_emb=True)
    g = cls(**params)
    assert params["upsampling_scales"] == [2, 4, 4, 5] and params["mid_channels"] == [192, 96, 48, 24]
    assert g.in_channels == 144 and g.upsampling_scales == [2, 4, 4, 5] and g.mid_channels == [192, 96, 48, 24]


def test_state_dict_layout_matches_reference(ref_keys):
    import harana.models as M
    g = M.FastSVCGenerator()
    sd = {k: list(v.shape) for k, v in g.state_dict().items()}
    assert len(sd) == 251 and sd == ref_keys["weight_norm"]
    assert list(sd) == list(ref_keys["weight_norm"])
    assert sum(p.numel() for p in g.parameters()) == 2751554
    g.remove_weight_norm()
    sd = {k: list(v.shape) for k, v in g.state_dict().items()}
    assert len(sd) == 170 and sd == ref_keys["plain"]
    assert sum(p.numel() for p in g.parameters()) == 2744353
    g.apply_weight_norm()
    assert len(g.state_dict()) == 251


def test_state_dict_round_trip_and_copy():
    import harana.models as M
    from svcc23_fastsvc_b200 import synthetic as syn
    g = M.FastSVCGenerator()
    p = syn.make_params(seed=5, weight_norm=True)
    g.load_state_dict({k: torch.from_numpy(v) for k, v in p.items()})
    for k, v in g.state_dict().items():
        assert torch.equal(v, torch.from_numpy(p[k]))
    g.remove_weight_norm()   # (deepcopy of old-style weight-normed modules is a torch limitation, same in the reference)
    g2 = copy.deepcopy(g)
    assert g2._handle is None and len(g2.state_dict()) == 170
    with pytest.raises(RuntimeError):
        g.load_state_dict({"bogus": torch.zeros(1)})


def test_constructor_errors_and_cpu_refusal():
    import harana.models as M
    with pytest.raises(ValueError):
        M.FastSVCGenerator(mid_channels=[8, 8], upsampling_scales=[2])
    g = M.FastSVCGenerator(in_channels=8, mid_channels=[8, 8], upsampling_scales=[2, 2], spk_emb_size=4)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        g(torch.zeros(1, 8, 3), torch.zeros(1, 1, 12), torch.zeros(1, 1, 12))
    with pytest.raises(ValueError):
        g.downsampling_loop(torch.zeros(1, 1, 4).cuda() if torch.cuda.is_available() else torch.zeros(1, 1, 4), 7, [])


def test_standalone_blocks_refuse_to_drop_gradients():
    """the block forwards return tensors without grad_fn; a grad-expecting call must raise, not train nothing
    (tacotron2.py:458-459 builds trainable FastSVCFiLMNet submodules)."""
    import harana.models as M
    film = M.FastSVCFiLMNet(8)
    with pytest.raises(RuntimeError, match="inference-only"):
        film(torch.zeros(1, 8, 4))
    with torch.no_grad(), pytest.raises(RuntimeError, match="no CPU fallback"):
        film(torch.zeros(1, 8, 4))
    down = M.FastSVCDownsampleNet(1, 8, 2)
    with pytest.raises(RuntimeError, match="inference-only"):
        down(torch.zeros(1, 1, 4))
    up = M.FastSVCUpsampleNet(8, 8, 2, 4)
    z = torch.zeros(1, 8, 8)
    with pytest.raises(RuntimeError, match="inference-only"):
        up(torch.zeros(1, 8, 4), (z, z), (z, z))


def test_forward_host_validates_shapes_before_touching_buffers():
    """fsvc_forward_host copies B*S / B*T floats from the host pointers -- every shape must be checked
    (and a (1, S) speaker expanded) before the call."""
    import harana.models as M
    g = M.FastSVCGenerator(in_channels=8, mid_channels=[8, 8], upsampling_scales=[2, 2], spk_emb_size=4)
    x, s, l = torch.zeros(2, 8, 3), torch.zeros(2, 1, 12), torch.zeros(2, 1, 12)
    with pytest.raises(ValueError, match="s and l must be"):
        g.forward_host(x, torch.zeros(2, 1, 11), l)
    with pytest.raises(ValueError, match="channels"):
        g.forward_host(torch.zeros(2, 7, 3), s, l)
    with pytest.raises(ValueError, match="spk_emb must be"):
        g.forward_host(x, s, l, torch.zeros(3, 4))
    B, frames, T, spk = g._check_inputs(x, s, l, torch.ones(1, 4))
    assert (B, frames, T) == (2, 3, 12) and tuple(spk.shape) == (2, 4)
    with pytest.raises(RuntimeError, match="no CPU fallback"):       # parameters on the CPU
        g.forward_host(x, s, l, torch.ones(1, 4))


def test_layers_semantics_cpu():
    from harana.layers import Squeeze2d, Stretch2d
    import numpy as np
    from conftest import load_golden
    gold = load_golden("layers")
    v = torch.arange(23, dtype=torch.float32).view(1, 1, 23)
    assert np.array_equal(Squeeze2d(5)(v).numpy(), gold["squeeze_23_5"])
    assert np.array_equal(Stretch2d(5, 1)(torch.arange(7, dtype=torch.float32).view(1, 1, 1, 7)).numpy(),
                          gold["stretch_7_5"])
