"""Host-side mirror of the reference's excitation helpers (harana/utils/features.py:21-216), backed by libfsvc.so.

``SignalGenerator`` keeps the reference's constructor, ``signal_types`` semantics and call signature
(``gen(f0) -> (B, n_types, T)``).  ``sinusoid`` runs the batched CUDA kernel ``fsvc_sine_excitation``; its Gaussian
noise is drawn with ``torch.randn`` on the input's device exactly where the reference draws it (features.py:194), so a
seeded run consumes the RNG stream the same way.  HOST tensors take the reference's own torch ops: that is the
data-pipeline use (the reference's ``Collater`` builds the excitation on the CPU inside DataLoader worker
processes, train_fastsvc.py:546), which north_star leaves in PyTorch; a CUDA tensor never falls back to it.
``F0Statistics`` is host-side numpy, as in the reference.
"""
import logging
import sys

import numpy as np
import torch

from . import abi

logger = logging.getLogger(__name__)


def _stream(device):
    return torch.cuda.current_stream(device).cuda_stream


def _require_cuda(t, what):
    if not t.is_cuda:
        raise RuntimeError(f"{what}: expected a CUDA tensor (libfsvc has no CPU fallback), got {t.device}")


class F0Statistics(object):
    """F0 statistics and the log-Gaussian mean/variance transformation (features.py:21-108)."""

    def estimate(self, f0list):
        f0s = np.concatenate([np.log(np.asarray(f0)[np.nonzero(f0)]) for f0 in f0list])
        return np.array([np.mean(f0s), np.std(f0s)])

    def convert(self, f0, orgf0stats, tarf0stats):
        f0 = np.asarray(f0)
        cvf0 = np.zeros(len(f0))
        nz = f0 > 0
        cvf0[nz] = np.exp((tarf0stats[1] / orgf0stats[1]) * (np.log(f0[nz]) - orgf0stats[0]) + tarf0stats[0])
        return cvf0


class SignalGenerator:
    """Input signal generator (features.py:145-216): "sine" (NSF excitation), "noise", "uv"."""

    def __init__(self, sample_rate=16000, hop_size=640, sine_amp=0.1, noise_amp=0.003,
                 signal_types=["sine", "noise"]):
        self.sample_rate = sample_rate
        self.hop_size = hop_size
        self.signal_types = signal_types
        self.sine_amp = sine_amp
        self.noise_amp = noise_amp
        for signal_type in signal_types:
            if signal_type not in ["noise", "sine", "uv"]:
                logger.info(f"{signal_type} is not supported type for generator input.")
                sys.exit(0)  # the reference exits here (features.py:174-176)
        logger.info(f"Use {signal_types} for generator input signals.")

    @torch.no_grad()
    def __call__(self, f0):
        signals = []
        for typ in self.signal_types:
            if "noise" == typ:
                signals.append(self.random_noise(f0))
            if "sine" == typ:
                signals.append(self.sinusoid(f0))
            if "uv" == typ:
                signals.append(self.vuv_binary(f0))
        return signals[0] if len(signals) == 1 else torch.cat(signals, axis=1)

    @torch.no_grad()
    def random_noise(self, f0):
        B, _, T = f0.size()
        return torch.randn((B, 1, T * self.hop_size), device=f0.device)

    @torch.no_grad()
    def sinusoid(self, f0, noise=None):
        """f0 (B, 1, T') -> NSF sine excitation (B, 1, T'*hop).  ``noise`` overrides the torch.randn draw (tests)."""
        if f0.dim() != 3 or f0.size(1) != 1:
            raise ValueError(f"f0 must be (B, 1, T'), got {tuple(f0.shape)}")
        B, _, T = f0.size()
        if not f0.is_cuda:
            return self._sinusoid_host(f0, noise)
        f0c = f0.to(torch.float32).contiguous()
        out = torch.empty((B, 1, T * self.hop_size), dtype=torch.float32, device=f0.device)
        nptr = 0
        if self.noise_amp > 0:
            if noise is None:
                noise = torch.randn((B, 1, T * self.hop_size), device=f0.device)
            noise = noise.to(torch.float32).contiguous()
            if noise.shape != out.shape or noise.device != f0.device:
                raise ValueError("noise must be (B, 1, T'*hop) on the device of f0")
            nptr = noise.data_ptr()
        with torch.cuda.device(f0.device):
            abi.check(abi.load().fsvc_sine_excitation(f0c.data_ptr(), nptr, out.data_ptr(), B, T, self.hop_size,
                                                      float(self.sample_rate), float(self.sine_amp),
                                                      float(self.noise_amp), _stream(f0.device)))
        return out

    def _sinusoid_host(self, f0, noise=None):
        """Host tensors (DataLoader workers): the reference's op sequence, features.py:188-195."""
        B, _, T = f0.size()
        n = T * self.hop_size
        vuv = torch.nn.functional.interpolate((f0 > 0) * torch.ones_like(f0), n)
        radious = (torch.nn.functional.interpolate(f0, n) / self.sample_rate) % 1
        sine = vuv * torch.sin(torch.cumsum(radious, dim=2) * 2 * np.pi) * self.sine_amp
        if self.noise_amp > 0:
            noise_amp = vuv * self.noise_amp + (1.0 - vuv) * self.noise_amp / 3.0
            if noise is None:
                noise = torch.randn((B, 1, n), device=f0.device)
            sine = sine + noise * noise_amp
        return sine

    @torch.no_grad()
    def vuv_binary(self, f0):
        return ((f0 > 0) * torch.ones_like(f0)).repeat_interleave(self.hop_size, dim=2)


def pcm16(wave):
    """Waveform tensor (any shape, CUDA fp32) -> int16 PCM as soundfile's "PCM_16" stores it: libsndfile's clipping
    conversion, floor(x * 32768) saturated to int16 (see csrc/excitation.cuh: restated, libsndfile is not in this image)."""
    _require_cuda(wave, "pcm16")
    w = wave.to(torch.float32).contiguous()
    out = torch.empty(w.shape, dtype=torch.int16, device=w.device)
    with torch.cuda.device(w.device):
        abi.check(abi.load().fsvc_pcm16(w.data_ptr(), out.data_ptr(), w.numel(), _stream(w.device)))
    return out
