"""Host-side mirrors of the reference's layer classes used by the FastSVC generator.

These exist so that (a) ``state_dict`` keys / shapes / weight-norm behaviour are
identical to the reference (the conv holders must be real ``nn.Conv1d`` /
``nn.Conv2d`` subclasses for ``torch.nn.utils.weight_norm`` to treat them the
same way) and (b) data-pipeline callers (``preprocess_fastsvc.py:71`` uses
``Stretch2d`` on a CPU tensor) keep working.  Inside the generator none of these
``forward`` methods run: the arithmetic is in libfsvc.so.

Reference: harana/layers/upsample.py:21-106, harana/layers/residual_block.py:27-48.
"""

import torch
import torch.nn.functional as F
from torch import nn


class Stretch2d(nn.Module):
    """Nearest-neighbour repeat of a (B, C, F, T) tensor (upsample.py:21-50)."""

    def __init__(self, x_scale, y_scale, mode="nearest"):
        super().__init__()
        self.x_scale, self.y_scale, self.mode = x_scale, y_scale, mode

    def forward(self, x):
        return F.interpolate(x, scale_factor=(self.y_scale, self.x_scale), mode=self.mode)


class Squeeze2d(nn.Module):
    """Nearest-neighbour decimation of a (B, C, T) tensor to int(T / scale) samples (upsample.py:53-74)."""

    def __init__(self, scale, mode="nearest"):
        super().__init__()
        self.scale, self.mode = scale, mode

    def forward(self, x):
        return F.interpolate(x, size=int(x.size(-1) / self.scale), mode=self.mode)


class Conv1d(nn.Conv1d):
    """``nn.Conv1d`` with kaiming-normal weights and zero bias (residual_block.py:27-38)."""

    def reset_parameters(self):
        nn.init.kaiming_normal_(self.weight, nonlinearity="relu")
        if self.bias is not None:
            nn.init.constant_(self.bias, 0.0)


class Conv1d1x1(Conv1d):
    """Pointwise conv parameter holder (residual_block.py:41-48)."""

    def __init__(self, in_channels, out_channels, bias=True):
        super().__init__(in_channels, out_channels, kernel_size=1, padding=0, dilation=1, bias=bias)


class Conv1d1x3(nn.Conv1d):
    """k=3 conv parameter holder, torch default init (upsample.py:76-83)."""

    def __init__(self, in_channels, out_channels, padding, dilation, bias=True):
        super().__init__(in_channels, out_channels, kernel_size=3, padding=padding, dilation=dilation, bias=bias)


class Conv2d1x3(nn.Conv2d):
    """(1,3) conv parameter holder on (B, C, 1, T), torch default init (upsample.py:99-106)."""

    def __init__(self, in_channels, out_channels, padding, dilation, bias=True):
        super().__init__(in_channels, out_channels, kernel_size=(1, 3), padding=padding, dilation=dilation,
                         bias=bias)


def effective_weight(conv):
    """The weight a conv holder currently stands for.  With old-style
    ``torch.nn.utils.weight_norm`` applied (fastsvc.py:354-362) that is
    ``g * v / ||v||`` (norm over all dims but 0), which the reference recomputes
    in a forward pre-hook on every call; our holders are never called, so it is
    computed here."""
    if hasattr(conv, "weight_g"):
        return torch._weight_norm(conv.weight_v, conv.weight_g, 0)
    return conv.weight
