"""Deterministic synthetic weights and inputs for the FastSVC generator.

Everything is drawn from ``numpy.random.RandomState`` (a frozen bit stream) so
that the golden fixtures under ``tests/golden`` -- which store only a seed plus
the reference's output -- can be regenerated bit-for-bit anywhere.  Shapes and
distributions follow SURVEY.md section 8(d).

No torch import here: numpy in, numpy out.
"""

from collections import OrderedDict

import numpy as np

YAML_CONFIG = dict(  # egs/svcc23/fastsvc1/conf/fastsvc.yaml:23-29
    in_channels=144,
    mid_channels=[192, 96, 48, 24],
    upsampling_scales=[2, 4, 4, 5],
    out_channels=1,
    spk_emb_size=512,
    use_spk_emb=True,
)


def hop_size(upsampling_scales):
    return int(np.prod(upsampling_scales))


def conv_specs(in_channels=144, mid_channels=(192, 96, 48, 24), upsampling_scales=(2, 4, 4, 5),
               out_channels=1, spk_emb_size=512, use_spk_emb=True):
    """Ordered list of (state_dict prefix, weight shape, kind) for every
    parametrised layer of ``FastSVCGenerator`` (reference fastsvc.py:261-301).
    kind is "conv2d", "conv1d" (weight-normalised in the reference) or
    "linear" (never weight-normalised, fastsvc.py:358)."""
    specs = []
    cin = in_channels
    for i, c in enumerate(mid_channels):
        p = f"upsampling_nets.{i}"
        specs.append((p + ".conv_first", (c, cin, 1, 3), "conv2d"))
        specs.append((p + ".upsample_block0.2", (c, c, 1, 3), "conv2d"))
        specs.append((p + ".conv_block1.1", (c, c, 1, 3), "conv2d"))
        specs.append((p + ".conv_block2.1", (c, c, 1, 3), "conv2d"))
        specs.append((p + ".conv_block3.1", (c, c, 1, 3), "conv2d"))
        specs.append((p + ".residual_block.1", (c, c, 1, 3), "conv2d"))
        if use_spk_emb:
            specs.append((p + ".emb_projector", (c, spk_emb_size), "linear"))
        cin = c
    for branch in ("downsampling_lft", "downsampling_sine"):
        cin = 1
        for i, c in enumerate(list(mid_channels)[::-1]):
            p = f"{branch}.{i}"
            specs.append((p + ".residual_block.0", (c, cin, 1), "conv1d"))
            specs.append((p + ".downsample_block.2", (c, cin, 3), "conv1d"))
            specs.append((p + ".downsample_block.4", (c, c, 3), "conv1d"))
            specs.append((p + ".downsample_block.6", (c, c, 3), "conv1d"))
            cin = c
    for branch in ("film_lft", "film_sine"):
        for i, c in enumerate(list(mid_channels)[::-1]):
            for sub in ("conv", "conv_scale", "conv_shift"):
                specs.append((f"{branch}.{i}.{sub}", (c, c, 3), "conv1d"))
    specs.append(("conv_last", (out_channels, mid_channels[-1], 1), "conv1d"))
    return specs


def make_params(config=None, seed=0, weight_norm=False, dtype=np.float32):
    """Random parameters keyed by the reference ``state_dict`` names.

    Weights/biases ~ U(-1/sqrt(fan_in), 1/sqrt(fan_in)) (torch's default conv
    init bound).  With ``weight_norm=True`` the convs are emitted in the
    weight-normalised form (``weight_g``, ``weight_v``) with a non-trivial gain.
    """
    cfg = dict(YAML_CONFIG if config is None else config)
    rs = np.random.RandomState(seed)
    out = OrderedDict()
    for prefix, shape, kind in conv_specs(**cfg):
        fan_in = int(np.prod(shape[1:]))
        bound = 1.0 / np.sqrt(fan_in)
        w = rs.uniform(-bound, bound, size=shape).astype(dtype)
        b = rs.uniform(-bound, bound, size=(shape[0],)).astype(dtype)
        if weight_norm and kind != "linear":
            axes = tuple(range(1, len(shape)))
            norm = np.sqrt((w.astype(np.float64) ** 2).sum(axis=axes, keepdims=True))
            gain = rs.uniform(0.5, 1.5, size=norm.shape)
            out[prefix + ".weight_g"] = (norm * gain).astype(dtype)
            out[prefix + ".weight_v"] = w
        else:
            out[prefix + ".weight"] = w
        out[prefix + ".bias"] = b
    return out


def make_f0(rs, B, frames):
    """Smooth-ish F0 contour in Hz, ~30 % unvoiced frames (f0 == 0)."""
    f0 = np.zeros((B, 1, frames), dtype=np.float32)
    for b in range(B):
        t = 0
        while t < frames:
            run = int(rs.randint(5, 40))
            voiced = rs.uniform() > 0.3
            if voiced:
                base = np.exp(rs.normal(np.log(220.0), 0.3))
                drift = np.cumsum(rs.normal(0.0, 0.01, size=run))
                f0[b, 0, t:t + run] = (base * np.exp(drift))[: max(0, min(run, frames - t))]
            t += run
    return f0


def sine_excitation(f0, rs, sample_rate=16000, hop=160, sine_amp=0.1, noise_amp=0.003):
    """NSF-style sine excitation, restating ``SignalGenerator.sinusoid``
    (reference harana/utils/features.py:178-197) in numpy: nearest-upsample
    F0, cumulative phase, sin, V/UV gate, additive Gaussian noise."""
    vuv = np.repeat((f0 > 0).astype(np.float32), hop, axis=-1)
    rad = np.mod(np.repeat(f0, hop, axis=-1) / np.float32(sample_rate), 1.0).astype(np.float32)
    phase = np.cumsum(rad, axis=2, dtype=np.float32)
    sine = vuv * np.sin(phase * np.float32(2 * np.pi)) * np.float32(sine_amp)
    if noise_amp > 0:
        namp = vuv * noise_amp + (1.0 - vuv) * noise_amp / 3.0
        sine = sine + rs.standard_normal(size=sine.shape).astype(np.float32) * namp
    return sine.astype(np.float32)


def make_inputs(B, frames, config=None, seed=1234, with_spk=True, lft_hop=64):
    """(ppg, sine, lft, spk) as float32 numpy arrays.

    ppg (B, Cin, T') ~ N(0,1) (PPGs are StandardScaler-normalised,
    normalize_fastsvc.py:131-135); sine (B,1,T) from a synthetic F0 contour;
    lft (B,1,T) piecewise-constant log-loudness over 64-sample hops
    (preprocess_fastsvc.py:60-75); spk (B,S) ~ N(0,1).
    """
    cfg = dict(YAML_CONFIG if config is None else config)
    hop = hop_size(cfg["upsampling_scales"])
    T = frames * hop
    rs = np.random.RandomState(seed)
    ppg = rs.standard_normal(size=(B, cfg["in_channels"], frames)).astype(np.float32)
    f0 = make_f0(rs, B, frames)
    sine = sine_excitation(f0, rs, hop=hop)
    nl = (T + lft_hop - 1) // lft_hop
    lft = np.clip(-4.0 + 2.0 * rs.standard_normal(size=(B, 1, nl)), -11.5, 3.0).astype(np.float32)
    lft = np.repeat(lft, lft_hop, axis=-1)[..., :T].copy()
    spk = rs.standard_normal(size=(B, cfg["spk_emb_size"])).astype(np.float32) if with_spk else None
    return ppg, sine, lft, spk
