"""The drop-in boundary against the reference's own scripts.

With this repo BEFORE the reference on ``sys.path`` (the "zero edit" route of INTEGRATION.md) every import the
reference's train / decode scripts make must resolve: names that live only in the reference (discriminators, HDF5 /
checkpoint helpers, losses, datasets) come from the reference, the generator classes come from this repo, and the
reference's ``Collater`` (which runs ``SignalGenerator`` on CPU tensors in DataLoader workers) works unchanged.

Runs in a subprocess (it rearranges ``sys.path`` / ``sys.modules``) and only where ``/root/reference`` exists
(the build container); the GPU box has no reference and runs the shim alone (tests/test_boundary.py)."""
import os
import subprocess
import sys
import textwrap

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"

pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "harana")), reason="needs /root/reference")

PRELUDE = f"""
import sys, types
REPO, REF = {REPO!r}, {REF!r}
sys.path[:] = [REPO, REF] + [p for p in sys.path if p not in ("", REPO, REF)]
# optional third-party modules the reference imports at module top and this image lacks (SURVEY.md 8c)
for name in ("h5py", "librosa", "kaldiio", "soundfile", "matplotlib", "matplotlib.pyplot"):
    sys.modules.setdefault(name, types.ModuleType(name))
sys.modules["matplotlib"].use = lambda *a, **k: None
sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
tk = types.ModuleType("tkinter"); tk.W = "w"; sys.modules.setdefault("tkinter", tk)
tb = types.ModuleType("tensorboardX"); tb.SummaryWriter = object; sys.modules.setdefault("tensorboardX", tb)
"""


def _run(body):
    code = PRELUDE + textwrap.dedent(body)
    env = dict(os.environ, PYTHONPATH="")
    res = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, cwd="/tmp",
                         timeout=600)
    assert res.returncode == 0, res.stdout + res.stderr
    return res.stdout


def test_reference_train_and_decode_imports_resolve_through_the_shim():
    out = _run("""
        import harana.models, harana.losses, harana.optimizers, harana.layers, harana.utils
        import svcc23_fastsvc_b200.generator as ours
        # generator classes: ours
        assert harana.models.__file__.startswith(REPO)
        assert getattr(harana.models, "FastSVCGenerator") is ours.FastSVCGenerator          # train_fastsvc.py:700-713
        from harana.models.fastsvc import FastSVCFiLMNet                                     # tacotron2.py:22
        assert FastSVCFiLMNet is ours.FastSVCFiLMNet
        # names that live only in the reference
        for name in ("MelGANMultiScaleDiscriminator", "MelGANDiscriminator", "HiFiGANMultiScaleMultiPeriodDiscriminator",
                     "HiFiGANPeriodDiscriminator", "HiFiGANScaleDiscriminator"):
            cls = getattr(harana.models, name)                                               # train_fastsvc.py:705
            assert cls.__module__.endswith("_reference_fastsvc"), cls.__module__
        from harana.utils import read_hdf5, load_model, write_hdf5, make_non_pad_mask        # train:38, decode:27-28
        assert read_hdf5.__module__ == "harana.utils.utils"
        from harana.losses import DiscriminatorAdversarialLoss, GeneratorAdversarialLoss, MultiResolutionSTFTLoss
        from harana.utils.features import SignalGenerator, F0Statistics                      # train:39, decode:29-30
        import svcc23_fastsvc_b200.features as feats
        assert SignalGenerator is feats.SignalGenerator
        from harana.layers import Stretch2d, Conv1d1x1, UpsampleNetwork, ResidualBlocks      # preprocess:35, hnusfgan
        import svcc23_fastsvc_b200.layers as L
        assert Stretch2d is L.Stretch2d and UpsampleNetwork.__module__.endswith("_reference_upsample")
        # the Tacotron2 model builds its FiLM nets from OUR class (tacotron2.py:458-459)
        import harana.models.tacotron2 as taco
        assert taco.FastSVCFiLMNet is ours.FastSVCFiLMNet and taco.__file__.startswith(REF)
        # the scripts themselves import
        import harana.bin.train_fastsvc as train
        import harana.bin.decode_fastsvc as decode
        assert train.__file__.startswith(REF) and decode.__file__.startswith(REF)
        assert train.SignalGenerator is feats.SignalGenerator
        # load_model's class lookup + state-dict load (utils.py:243-280) builds OUR generator from a checkpoint
        import torch, yaml, tempfile, os
        g = ours.FastSVCGenerator()
        d = tempfile.mkdtemp()
        torch.save({"model": {"generator": g.state_dict()}}, os.path.join(d, "checkpoint-1steps.pkl"))
        cfg = {"generator_type": "FastSVCGenerator", "generator_params": dict(
            in_channels=144, out_channels=1, mid_channels=[192, 96, 48, 24], upsampling_scales=[2, 4, 4, 5],
            spk_emb_size=512, use_spk_emb=True)}
        m = load_model(os.path.join(d, "checkpoint-1steps.pkl"), config=cfg)
        assert type(m) is ours.FastSVCGenerator
        for k, v in g.state_dict().items():
            assert torch.equal(v, m.state_dict()[k])
        print("OK")
    """)
    assert out.strip().endswith("OK")


def test_reference_collater_runs_on_cpu_with_the_shimmed_signal_generator():
    out = _run("""
        import numpy as np, torch
        import harana.bin.train_fastsvc as train
        import svcc23_fastsvc_b200.features as feats
        col = train.Collater(batch_length=8192, sample_rate=16000, hop_size=160, sine_amp=0.1, noise_amp=0.003,
                             signal_types=["sine"], use_spk_emb=True)                       # train_fastsvc.py:660-671
        assert isinstance(col.signal_generator, feats.SignalGenerator) and col.batch_length == 8160
        rs = np.random.RandomState(0)
        items = []
        for i in range(3):
            frames = 80 + 7 * i
            f0 = np.exp(np.log(200.0) + 0.2 * rs.randn(frames)) * (rs.rand(frames) > 0.3)
            items.append((rs.randn(frames * 160), f0, rs.randn(frames, 144), rs.randn(frames * 160), rs.randn(512, 1)))
        np.random.seed(1); torch.manual_seed(2)
        (ppg, sine, lft, emb), y = col(items)
        assert ppg.shape == (3, 144, 51) and sine.shape == (3, 1, 8160) and lft.shape == (3, 1, 8160)
        assert emb.shape == (3, 512) and y.shape == (3, 1, 8160) and not sine.is_cuda
        # same crop + same draw through the REFERENCE's SignalGenerator: bit-identical excitation
        import harana.utils.features as shim
        ref_gen = shim._dropin.load_shadowed("harana.utils", shim.__file__, "features").SignalGenerator(
            sample_rate=16000, hop_size=160, sine_amp=0.1, noise_amp=0.003, signal_types=["sine"])
        col.signal_generator = ref_gen
        np.random.seed(1); torch.manual_seed(2)
        (_, sine_ref, _, _), _ = col(items)
        assert torch.equal(sine, sine_ref)
        print("OK")
    """)
    assert out.strip().endswith("OK")
