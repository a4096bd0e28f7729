"""CPU baseline port (PyTorch ops) of the reference FastSVC generator forward.

TEST / BASELINE INFRASTRUCTURE ONLY -- never imported by the product path.
Used by ``bench.py`` (``cpu_baseline`` leg and ``--impl reference``) and by
tests as a second checker.  ``/root/reference`` does not exist on the GPU box,
so this port stands in for the reference there; it issues the *same sequence of
stock ``torch.nn.functional`` ops as the reference does*, including the
reference's re-computation of the conditioning chain for every stage
(fastsvc.py:322-326, 334-340) and its Conv2d-on-(B,C,1,T) formulation, so that
its CPU (oneDNN) and eager-CUDA (cuDNN) timings are the reference's timings.

Parity status: PINNED against the imported reference via ``tests/golden``
(see ``tests/golden/make_golden.py`` and ``tests/test_oracle.py``).

Parameters: flat dict of tensors keyed by the reference's ``state_dict`` names
(weight-normalised or plain).
"""

import torch
import torch.nn.functional as F


def effective_weight(params, prefix):
    """``torch.nn.utils.weight_norm`` pre-hook (applied at fastsvc.py:354-362):
    w = g * v / ||v|| over all dims but 0."""
    if prefix + ".weight" in params:
        return params[prefix + ".weight"]
    g, v = params[prefix + ".weight_g"], params[prefix + ".weight_v"]
    return g * v / v.flatten(1).norm(dim=1).view(-1, *([1] * (v.dim() - 1)))


def resolve_weights(params):
    """Fold weight norm once (what ``remove_weight_norm`` does in decode,
    decode_fastsvc.py:142)."""
    out = {}
    for k, t in params.items():
        if k.endswith(".weight_g"):
            p = k[: -len(".weight_g")]
            out[p + ".weight"] = effective_weight(params, p)
        elif k.endswith(".weight_v"):
            continue
        else:
            out[k] = t
    return out


def _lrelu(x):
    return F.leaky_relu(x, 0.2)


def _squeeze(x, scale):
    # layers/upsample.py:64-74
    return F.interpolate(x, size=int(x.size(-1) / scale), mode="nearest")


def _stretch(x, scale):
    # layers/upsample.py:38-50 (x is (B,C,1,T))
    return F.interpolate(x, scale_factor=(1, scale), mode="nearest")


def _c1(p, prefix, x, dilation, padding):
    return F.conv1d(x, p[prefix + ".weight"], p[prefix + ".bias"], dilation=dilation, padding=padding)


def _c2(p, prefix, x, dilation):
    # Conv2d1x3(..., (0, d), d): layers/upsample.py:99-106
    return F.conv2d(x, p[prefix + ".weight"], p[prefix + ".bias"], dilation=dilation, padding=(0, dilation))


def downsample_net(p, prefix, x, scale):
    """fastsvc.py:180-193."""
    r = _squeeze(_c1(p, prefix + ".residual_block.0", x, 1, 0), scale)
    h = _lrelu(_squeeze(x, scale))
    h = _c1(p, prefix + ".downsample_block.2", h, 1, 1)
    h = _c1(p, prefix + ".downsample_block.4", _lrelu(h), 2, 2)
    h = _c1(p, prefix + ".downsample_block.6", _lrelu(h), 4, 4)
    return h + r


def film_net(p, prefix, x):
    """fastsvc.py:220-232."""
    h = _lrelu(_c1(p, prefix + ".conv", x, 1, 1))
    return _c1(p, prefix + ".conv_scale", h, 1, 1), _c1(p, prefix + ".conv_shift", h, 1, 1)


def feature_affine(p, prefix, x, sine, lft, spk_emb):
    """fastsvc.py:115-140 (x is (B,C,1,T))."""
    s_scale, s_shift = sine
    l_scale, l_shift = lft
    scale = s_scale + l_scale
    shift = s_shift.unsqueeze(2) + l_shift.unsqueeze(2)
    x = torch.mul(scale.unsqueeze(2), x)
    x = x + shift
    if spk_emb is not None:
        e = F.linear(F.normalize(spk_emb), p[prefix + ".emb_projector.weight"],
                     p[prefix + ".emb_projector.bias"]).unsqueeze(2).unsqueeze(3)
        x = F.instance_norm(x, eps=1e-5)
        x = x + e
    return x


def upsample_net(p, prefix, x, sine, lft, scale, spk_emb):
    """fastsvc.py:80-113."""
    x = x.unsqueeze(2)
    x = _c2(p, prefix + ".conv_first", x, 1)
    xr = _c2(p, prefix + ".residual_block.1", _stretch(x, scale), 1)
    x = _lrelu(_c2(p, prefix + ".upsample_block0.2", _stretch(_lrelu(x), scale), 1))
    x = feature_affine(p, prefix, x, sine, lft, spk_emb)
    x = _c2(p, prefix + ".conv_block1.1", _lrelu(x), 3)
    x_ = x + xr
    x = feature_affine(p, prefix, x_, sine, lft, spk_emb)
    x = _c2(p, prefix + ".conv_block2.1", _lrelu(x), 9)
    x = feature_affine(p, prefix, x, sine, lft, spk_emb)
    x = _c2(p, prefix + ".conv_block3.1", _lrelu(x), 27)
    x = x + x_
    return x.squeeze(2)


def downsampling_scales(upsampling_scales):
    d = list(upsampling_scales)[::-1]
    d.pop()
    d.insert(0, 1)
    return d


def generator_forward(params, x, s, l, spk_emb=None, upsampling_scales=(2, 4, 4, 5), recompute=True):
    """fastsvc.py:305-340.  ``recompute=True`` re-runs the conditioning chain
    per stage exactly like ``downsampling_loop`` does (reference timing);
    ``recompute=False`` runs it once (same values)."""
    p = resolve_weights(params)
    n = len(upsampling_scales)
    dscales = downsampling_scales(upsampling_scales)

    def chain(sig, name, upto):
        for i in range(upto + 1):
            sig = downsample_net(p, f"{name}.{i}", sig, dscales[i])
        return sig

    cache = {}
    if not recompute:
        for name, sig in (("downsampling_lft", l), ("downsampling_sine", s)):
            h = sig
            for i in range(n):
                h = downsample_net(p, f"{name}.{i}", h, dscales[i])
                cache[(name, i)] = h
    for idx in range(n):
        didx = n - idx - 1
        if recompute:
            ld = chain(l, "downsampling_lft", didx)
            sd = chain(s, "downsampling_sine", didx)
        else:
            ld, sd = cache[("downsampling_lft", didx)], cache[("downsampling_sine", didx)]
        lft = film_net(p, f"film_lft.{didx}", ld)
        sine = film_net(p, f"film_sine.{didx}", sd)
        x = upsample_net(p, f"upsampling_nets.{idx}", x, sine, lft, upsampling_scales[idx], spk_emb)
    return _c1(p, "conv_last", x, 1, 0)
