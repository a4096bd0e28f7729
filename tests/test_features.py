"""Excitation path: oracle vs reference-generated golden vectors (CPU) and the CUDA kernel vs both (GPU).

Golden files: tests/golden/sine_*.npz, f0_convert.npz -- generated from the reference's harana/utils/features.py by
tests/golden/make_golden_features.py."""
import numpy as np
import pytest
import torch

from conftest import load_golden
from oracle import features_numpy as fo

SINE_CASES = ["sine_b2_f20", "sine_b3_f100", "sine_b1_f500", "sine_nonoise_b2_f33", "sine_hop64_24k_b2_f50"]
TOL_ORACLE = 5e-8   # numpy vs torch CPU sin differ by <= 1 ulp of a value <= 0.25
TOL_GPU = 2e-6      # sinf (<= 2 ulp at arguments up to ~7e3 rad) * amplitude; the phase itself is bit-identical


def _meta(g):
    B, frames, hop, sr, samp, namp = g["meta"]
    return int(B), int(frames), int(hop), float(sr), float(samp), float(namp)


@pytest.mark.parametrize("name", SINE_CASES)
def test_oracle_matches_reference_golden(name):
    g = load_golden(name)
    B, frames, hop, sr, samp, namp = _meta(g)
    out = fo.sinusoid(g["f0"], g["noise"], sr, hop, samp, namp)
    assert out.shape == (B, 1, frames * hop) and out.dtype == np.float32
    assert np.abs(out - g["out"]).max() <= TOL_ORACLE
    assert np.array_equal(fo.vuv_binary(g["f0"], hop), g["uv"])


def test_f0_convert_matches_reference_golden():
    from svcc23_fastsvc_b200.features import F0Statistics
    g = load_golden("f0_convert")
    assert np.array_equal(fo.f0_convert(g["f0"], g["src"], g["trg"]), g["out"])
    assert np.array_equal(F0Statistics().convert(g["f0"], g["src"], g["trg"]), g["out"])
    st = F0Statistics().estimate([g["f0"][:100], g["f0"][100:]])
    lf = np.log(g["f0"][g["f0"] > 0])
    assert np.allclose(st, [lf.mean(), lf.std()])


@pytest.mark.parametrize("name", SINE_CASES)
def test_signal_generator_host_tensors_match_reference_golden(name):
    """Host tensors = the data-pipeline use (the reference's Collater runs SignalGenerator on the CPU in DataLoader
    workers, train_fastsvc.py:546): the reference's torch ops, bit for bit; the seeded draw is the reference's."""
    from harana.utils.features import SignalGenerator
    g = load_golden(name)
    B, frames, hop, sr, samp, namp = _meta(g)
    gen = SignalGenerator(sample_rate=sr, hop_size=hop, sine_amp=samp, noise_amp=namp, signal_types=["sine"])
    out = gen.sinusoid(torch.from_numpy(g["f0"]), noise=torch.from_numpy(g["noise"]))
    assert np.array_equal(out.numpy(), g["out"])
    seed = {"sine_b2_f20": 11, "sine_b3_f100": 12, "sine_b1_f500": 13, "sine_nonoise_b2_f33": 14,
            "sine_hop64_24k_b2_f50": 15}[name]
    torch.manual_seed(seed)
    assert np.array_equal(gen(torch.from_numpy(g["f0"])).numpy(), g["out"])
    uv = SignalGenerator(sample_rate=sr, hop_size=hop, signal_types=["uv"])(torch.from_numpy(g["f0"]))
    assert np.array_equal(uv.numpy(), g["uv"])


def test_pcm16_oracle_known_answers():
    """libsndfile f2les_clip_array with clipping on (what soundfile.write(..., "PCM_16") runs): top 16 bits of
    lrintf(x * 2^31), i.e. floor(x * 32768) saturated."""
    x = np.array([0, 1, -1, 0.5 / 32768, -0.5 / 32768, 1 / 32768, -1 / 32768, 32767 / 32768, 0.99999, -0.99999, 1.5,
                  -1.5, 1e-9, -1e-9, 0.25, -0.25], dtype=np.float32)
    want = [0, 32767, -32768, 0, -1, 1, -1, 32767, 32767, -32768, 32767, -32768, 0, -1, 8192, -8192]
    assert fo.pcm16(x).tolist() == want
    rs = np.random.RandomState(1)
    y = rs.uniform(-1.0, 1.0, 10000).astype(np.float32)
    assert np.array_equal(fo.pcm16(y), np.floor(y.astype(np.float64) * 32768.0).clip(-32768, 32767).astype(np.int16))


@pytest.mark.gpu
@pytest.mark.parametrize("name", SINE_CASES)
def test_cuda_excitation_matches_golden(name):
    from harana.utils.features import SignalGenerator
    g = load_golden(name)
    B, frames, hop, sr, samp, namp = _meta(g)
    dev = torch.device("cuda:0")
    gen = SignalGenerator(sample_rate=sr, hop_size=hop, sine_amp=samp, noise_amp=namp, signal_types=["sine"])
    out = gen.sinusoid(torch.from_numpy(g["f0"]).to(dev), noise=torch.from_numpy(g["noise"]).to(dev))
    assert out.shape == (B, 1, frames * hop) and out.dtype == torch.float32 and out.is_cuda
    got = out.cpu().numpy()
    assert np.abs(got - g["out"]).max() <= TOL_GPU
    assert np.abs(got - fo.sinusoid(g["f0"], g["noise"], sr, hop, samp, namp)).max() <= TOL_GPU
    uv = SignalGenerator(sample_rate=sr, hop_size=hop, signal_types=["uv"])(torch.from_numpy(g["f0"]).to(dev))
    assert np.array_equal(uv.cpu().numpy(), g["uv"])


@pytest.mark.gpu
def test_cuda_excitation_draws_noise_like_the_reference():
    """The seeded torch.randn draw happens inside the call, with the reference's shape (features.py:194)."""
    from harana.utils.features import SignalGenerator
    dev = torch.device("cuda:0")
    f0 = (200.0 + 50.0 * torch.rand(4, 1, 37, device=dev)) * (torch.rand(4, 1, 37, device=dev) > 0.3)
    gen = SignalGenerator(sample_rate=16000, hop_size=160, signal_types=["sine", "uv"])
    torch.manual_seed(7)
    a = gen(f0)
    torch.manual_seed(7)
    noise = torch.randn((4, 1, 37 * 160), device=dev)
    b = gen.sinusoid(f0, noise=noise)
    assert a.shape == (4, 2, 37 * 160)
    assert torch.equal(a[:, :1], b)
    # long, batched, full-size: unvoiced samples carry noise only; voiced amplitude bounded by sine_amp + noise tail
    f0 = torch.full((32, 1, 500), 220.0, device=dev)
    f0[:, :, 100:200] = 0.0
    s = gen.sinusoid(f0, noise=torch.zeros(32, 1, 80000, device=dev))
    assert float(s[:, :, 16000:32000].abs().max()) == 0.0
    assert abs(float(s.abs().max()) - 0.1) < 1e-6
    # linear phase at constant f0: zero crossings every fs / (2 f0) samples in the first voiced run
    x = s[0, 0, :16000].cpu().numpy().astype(np.float64)
    ref = 0.1 * np.sin(2 * np.pi * 220.0 / 16000.0 * np.arange(1, 16001))
    assert np.abs(x - ref).max() < 2e-4


@pytest.mark.gpu
def test_pcm16_matches_soundfile_rule():
    from svcc23_fastsvc_b200.features import pcm16
    rs = np.random.RandomState(0)
    x = np.concatenate([rs.uniform(-1.2, 1.2, 100000), [0.0, 1.0, -1.0, 0.5 / 32768, -0.5 / 32768, 1.0 / 32768, -1.0 / 32768, 32767.0 / 32768, 0.99999, -0.99999,
                             1e-9, -1e-9]]).astype(np.float32)
    y = pcm16(torch.from_numpy(x).cuda()).cpu().numpy()
    want = fo.pcm16(x)
    assert y.dtype == np.int16 and np.array_equal(y, want)
