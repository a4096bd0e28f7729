"""CPU oracle (numpy) for the FastSVC generator forward pass.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import it, and only as the checker or the timed
baseline.  The product path (``svcc23_fastsvc_b200`` / ``harana``) never
imports this module and has no CPU fallback.

This file restates, function by function, what the reference computes.  The
arithmetic of the reference lives in PyTorch ATen (``nn.Conv1d``/``Conv2d``,
``F.interpolate``, ``nn.InstanceNorm2d``, ``nn.Linear``, ``F.normalize``,
``nn.LeakyReLU``; reference pins torch==1.12.0 in ``setup.py:27``, this image
has torch 2.11) -- third-party code that is not under ``/root/reference`` -- so
each function below restates the *published* semantics of the ATen op at the
reference's call site and cites that call site.

Parity status: PINNED.  The reference ships no tests or golden vectors
(SURVEY.md section 4), so the oracle is pinned against outputs of the reference
itself, imported unmodified in the build container by
``tests/golden/make_golden.py``; the resulting fixtures are committed under
``tests/golden/`` and checked by ``tests/test_oracle.py``.

All functions take/return ``numpy`` arrays in the reference's layouts
(``(B, C, T)``), compute in the dtype of the inputs (fp32 or fp64), and take
parameters as a flat ``dict`` keyed by the reference's ``state_dict`` names
(either weight-normalised ``*.weight_g`` / ``*.weight_v`` or plain
``*.weight``).
"""

import numpy as np

LRELU_SLOPE = 0.2  # nn.LeakyReLU(0.2): fastsvc.py:58,61,64,67,70,172-176,212
IN_EPS = 1e-5  # nn.InstanceNorm2d default eps: fastsvc.py:76


# --------------------------------------------------------------------------
# elementary ops
# --------------------------------------------------------------------------
def leaky_relu(x, slope=LRELU_SLOPE):
    """``nn.LeakyReLU(0.2)`` (fastsvc.py:58): x if x >= 0 else slope * x."""
    return np.where(x >= 0, x, x * np.asarray(slope, dtype=x.dtype))


def conv1d(x, weight, bias=None, dilation=1, padding=0):
    """Zero-padded, stride-1 cross-correlation, as ``nn.Conv1d`` / ``nn.Conv2d``
    with a (1,k) kernel on a height-1 image compute it.

    Call sites: ``Conv1d1x1`` (layers/residual_block.py:41-48), ``Conv1d1x3``
    (layers/upsample.py:76-83), ``Conv2d1x3`` (layers/upsample.py:99-106; the
    input is unsqueezed to (B,C,1,T) at fastsvc.py:92, so the 2-D conv is
    exactly this 1-D conv with weight[:, :, 0, :]).

    x: (B, Cin, T); weight: (Cout, Cin, k) or (Cout, Cin, 1, k); bias: (Cout,).
    """
    if weight.ndim == 4:
        weight = weight[:, :, 0, :]
    B, cin, T = x.shape
    cout, cin_w, k = weight.shape
    assert cin == cin_w, (cin, cin_w)
    t_out = T + 2 * padding - dilation * (k - 1)
    xp = np.zeros((B, cin, T + 2 * padding), dtype=x.dtype)
    xp[:, :, padding:padding + T] = x
    out = np.zeros((B, cout, t_out), dtype=x.dtype)
    for j in range(k):
        out += np.matmul(weight[None, :, :, j], xp[:, :, j * dilation:j * dilation + t_out])
    if bias is not None:
        out += bias[None, :, None]
    return out


def stretch(x, scale):
    """``Stretch2d(scale, 1)`` (layers/upsample.py:38-50):
    ``F.interpolate(scale_factor=(1, scale), mode="nearest")`` on (B,C,1,T).
    With an integer scale factor ATen's nearest index is floor(i / scale),
    i.e. every sample is repeated ``scale`` times."""
    return np.repeat(x, int(scale), axis=-1)


def squeeze(x, scale):
    """``Squeeze2d(scale)`` (layers/upsample.py:64-74):
    ``F.interpolate(size=int(T / scale), mode="nearest")``.
    ATen nearest: src = min(floor(dst * float32(T / size)), T - 1).  When
    ``T % scale == 0`` this is ``x[..., ::scale]``."""
    T = x.shape[-1]
    size = int(T / scale)
    ratio = np.float32(T) / np.float32(size)
    idx = np.minimum(np.floor(np.arange(size, dtype=np.float32) * ratio).astype(np.int64), T - 1)
    return x[..., idx]


def instance_norm(x, eps=IN_EPS):
    """``nn.InstanceNorm2d(C)`` (fastsvc.py:76,138): no affine, no running
    stats (same in train and eval): per-(b, c) mean and *biased* variance over
    the whole time axis."""
    mean = x.mean(axis=-1, keepdims=True)
    var = ((x - mean) ** 2).mean(axis=-1, keepdims=True)
    return (x - mean) / np.sqrt(var + np.asarray(eps, dtype=x.dtype))


def l2_normalize(x, eps=1e-12):
    """``F.normalize(spk_emb)`` (fastsvc.py:136): x / max(||x||_2, eps) over dim 1."""
    n = np.sqrt((x * x).sum(axis=1, keepdims=True))
    return x / np.maximum(n, np.asarray(eps, dtype=x.dtype))


def linear(x, weight, bias):
    """``nn.Linear`` (fastsvc.py:78,136): x @ W^T + b."""
    return x @ weight.T + bias[None, :]


def effective_weight(params, prefix):
    """Weight of the conv at ``prefix``.  With ``torch.nn.utils.weight_norm``
    applied (fastsvc.py:354-362) the state dict holds ``weight_g`` (Cout,1,1[,1])
    and ``weight_v``; the forward pre-hook recomputes
    ``w = g * v / ||v||_2`` with the norm over every dim but 0."""
    if prefix + ".weight" in params:
        return params[prefix + ".weight"]
    g = params[prefix + ".weight_g"]
    v = params[prefix + ".weight_v"]
    axes = tuple(range(1, v.ndim))
    norm = np.sqrt((v * v).sum(axis=axes, keepdims=True))
    return g * v / norm


def _conv(params, prefix, x, dilation, padding):
    return conv1d(x, effective_weight(params, prefix), params[prefix + ".bias"], dilation, padding)


# --------------------------------------------------------------------------
# blocks
# --------------------------------------------------------------------------
def downsample_net(params, prefix, x, scale):
    """``FastSVCDownsampleNet.forward`` (fastsvc.py:180-193; ctor 146-178).

    r = Squeeze(Conv1x1(x));
    d = Conv3_d4(lrelu(Conv3_d2(lrelu(Conv3_d1(lrelu(Squeeze(x)))))));  out = d + r.
    """
    r = squeeze(_conv(params, prefix + ".residual_block.0", x, 1, 0), scale)
    h = leaky_relu(squeeze(x, scale))
    h = _conv(params, prefix + ".downsample_block.2", h, 1, 1)
    h = _conv(params, prefix + ".downsample_block.4", leaky_relu(h), 2, 2)
    h = _conv(params, prefix + ".downsample_block.6", leaky_relu(h), 4, 4)
    return h + r


def film_net(params, prefix, x):
    """``FastSVCFiLMNet.forward`` (fastsvc.py:220-232): h = lrelu(conv(x));
    scale = conv_scale(h); shift = conv_shift(h)."""
    h = leaky_relu(_conv(params, prefix + ".conv", x, 1, 1))
    return _conv(params, prefix + ".conv_scale", h, 1, 1), _conv(params, prefix + ".conv_shift", h, 1, 1)


def feature_affine(params, prefix, x, sine, lft, spk_emb=None):
    """``FastSVCUpsampleNet._feature_affine`` (fastsvc.py:115-140).

    x * (s_scale + l_scale) + (s_shift + l_shift); when a speaker embedding is
    given: InstanceNorm, then + Linear(normalize(spk_emb)) broadcast over time.
    """
    s_scale, s_shift = sine
    l_scale, l_shift = lft
    x = (s_scale + l_scale) * x
    x = x + (s_shift + l_shift)
    if spk_emb is not None:
        e = linear(l2_normalize(spk_emb), params[prefix + ".emb_projector.weight"],
                   params[prefix + ".emb_projector.bias"])
        x = instance_norm(x) + e[:, :, None]
    return x


def upsample_net(params, prefix, x, sine, lft, scale, spk_emb=None):
    """``FastSVCUpsampleNet.forward`` (fastsvc.py:80-113; ctor 37-78)."""
    x = _conv(params, prefix + ".conv_first", x, 1, 1)                      # :93
    xr = _conv(params, prefix + ".residual_block.1", stretch(x, scale), 1, 1)  # :94, :72-75
    x = leaky_relu(_conv(params, prefix + ".upsample_block0.2",
                         stretch(leaky_relu(x), scale), 1, 1))              # :97, :57-62
    x = feature_affine(params, prefix, x, sine, lft, spk_emb)              # :98
    x = _conv(params, prefix + ".conv_block1.1", leaky_relu(x), 3, 3)       # :99
    x_ = x + xr                                                             # :102
    x = feature_affine(params, prefix, x_, sine, lft, spk_emb)             # :105
    x = _conv(params, prefix + ".conv_block2.1", leaky_relu(x), 9, 9)       # :106
    x = feature_affine(params, prefix, x, sine, lft, spk_emb)              # :107
    x = _conv(params, prefix + ".conv_block3.1", leaky_relu(x), 27, 27)     # :108
    return x + x_                                                           # :111


def downsampling_scales(upsampling_scales):
    """fastsvc.py:270-272: reverse, drop the last, put 1 in front."""
    d = list(upsampling_scales)[::-1]
    d.pop()
    d.insert(0, 1)
    return d


def generator_forward(params, x, s, l, spk_emb=None, upsampling_scales=(2, 4, 4, 5)):
    """``FastSVCGenerator.forward`` (fastsvc.py:305-332).

    The reference re-runs the conditioning chain from scratch for every stage
    (``downsampling_loop``, fastsvc.py:334-340); the chain has no state, so
    running it once and keeping every level's output is the same function.

    x: (B, Cin, T'), s, l: (B, 1, T' * prod(scales)), spk_emb: (B, S) or None.
    Returns (B, out_channels, T).
    """
    n = len(upsampling_scales)
    dscales = downsampling_scales(upsampling_scales)
    lft_levels, sine_levels = [], []
    hl, hs = l, s
    for i in range(n):
        hl = downsample_net(params, f"downsampling_lft.{i}", hl, dscales[i])
        hs = downsample_net(params, f"downsampling_sine.{i}", hs, dscales[i])
        lft_levels.append(hl)
        sine_levels.append(hs)
    for idx in range(n):
        didx = n - idx - 1
        lft = film_net(params, f"film_lft.{didx}", lft_levels[didx])
        sine = film_net(params, f"film_sine.{didx}", sine_levels[didx])
        x = upsample_net(params, f"upsampling_nets.{idx}", x, sine, lft,
                         upsampling_scales[idx], spk_emb)
    return _conv(params, "conv_last", x, 1, 0)                             # :330
