"""Host-PyTorch critics and losses that DRIVE the training benchmarks (BASELINE configs[2]/[3]); not product code.

north_star keeps the GAN losses and discriminators in host PyTorch, and the reference's own classes cannot travel to
the GPU box (/root/reference does not exist there), so this file restates -- in stock torch.nn, for the benchmark only
-- the pieces ``Trainer._train_step`` touches besides the generator:

  * HiFiGAN multi-scale + multi-period discriminator, defaults of harana/models/fastsvc.py:1055-1143 (scale critic
    :818-975, period critic :631-760).  Quirks kept: the scale critics never get weight / spectral norm (the reference's
    ``apply_weight_norm`` tests ``isinstance(m, nn.Conv2d)`` on Conv1d layers, :957-975) and the period critic's output
    conv has kernel ``(kernel_sizes[1] - 1, 1)`` (:689-695).
  * multi-resolution STFT loss, harana/losses/stft_loss.py:21-180, with the recipe's six resolutions
    (egs/svcc23/fastsvc1/conf/fastsvc.yaml:57-61).
  * LSGAN ("mse") adversarial losses, harana/losses/adversarial_loss.py:16-127 with average_by_discriminators=True.
"""
import torch
import torch.nn.functional as F
from torch import nn

STFT_PARAMS = dict(fft_sizes=[2048, 1024, 512, 256, 128, 64], hop_sizes=[512, 256, 128, 64, 32, 16],
                   win_lengths=[2048, 1024, 512, 256, 128, 64])


class ScaleCritic(nn.Module):
    def __init__(self, kernel_sizes=(15, 41, 5, 3), channels=128, max_channels=1024, max_groups=16,
                 downsample_scales=(2, 2, 4, 4, 1), slope=0.1):
        super().__init__()
        self.layers = nn.ModuleList([nn.Sequential(
            nn.Conv1d(1, channels, kernel_sizes[0], padding=(kernel_sizes[0] - 1) // 2), nn.LeakyReLU(slope))])
        in_chs, out_chs, groups = channels, channels, 4
        for s in downsample_scales:
            self.layers.append(nn.Sequential(
                nn.Conv1d(in_chs, out_chs, kernel_sizes[1], stride=s, padding=(kernel_sizes[1] - 1) // 2,
                          groups=groups), nn.LeakyReLU(slope)))
            in_chs = out_chs
            out_chs = min(in_chs * 2, max_channels)
            groups = min(groups * 4, max_groups)
        out_chs = min(in_chs * 2, max_channels)
        self.layers.append(nn.Sequential(
            nn.Conv1d(in_chs, out_chs, kernel_sizes[2], padding=(kernel_sizes[2] - 1) // 2), nn.LeakyReLU(slope)))
        self.last_layer = nn.Conv1d(out_chs, 1, kernel_sizes[3], padding=(kernel_sizes[3] - 1) // 2)

    def forward(self, x):
        for f in self.layers:
            x = f(x)
        return self.last_layer(x)


class PeriodCritic(nn.Module):
    def __init__(self, period, kernel_sizes=(5, 3), channels=32, downsample_scales=(3, 3, 3, 3, 1),
                 max_channels=1024, slope=0.1):
        super().__init__()
        self.period = period
        self.convs = nn.ModuleList()
        in_chs, out_chs = 1, channels
        for s in downsample_scales:
            self.convs.append(nn.Sequential(
                nn.utils.weight_norm(nn.Conv2d(in_chs, out_chs, (kernel_sizes[0], 1), (s, 1),
                                               padding=((kernel_sizes[0] - 1) // 2, 0))), nn.LeakyReLU(slope)))
            in_chs = out_chs
            out_chs = min(out_chs * 4, max_channels)
        self.output_conv = nn.utils.weight_norm(nn.Conv2d(out_chs, 1, (kernel_sizes[1] - 1, 1), 1,
                                                          padding=((kernel_sizes[1] - 1) // 2, 0)))

    def forward(self, x):
        b, c, t = x.shape
        if t % self.period != 0:
            n_pad = self.period - (t % self.period)
            x = F.pad(x, (0, n_pad), "reflect")
            t += n_pad
        x = x.view(b, c, t // self.period, self.period)
        for f in self.convs:
            x = f(x)
        return torch.flatten(self.output_conv(x), 1, -1)


class MultiScaleMultiPeriodCritic(nn.Module):
    """HiFiGANMultiScaleMultiPeriodDiscriminator() with its default arguments (70.7 M parameters)."""

    def __init__(self, scales=3, periods=(2, 3, 5, 7, 11)):
        super().__init__()
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore", FutureWarning)
            self.msd = nn.ModuleList([ScaleCritic() for _ in range(scales)])
            self.mpd = nn.ModuleList([PeriodCritic(p) for p in periods])
        self.pooling = nn.AvgPool1d(kernel_size=4, stride=2, padding=2)

    def forward(self, x):
        outs = []
        xs = x
        for f in self.msd:
            outs.append(f(xs))
            xs = self.pooling(xs)
        for f in self.mpd:
            outs.append(f(x))
        return outs


def _stft_mag(x, fft_size, hop, win_length, window):
    s = torch.stft(x, fft_size, hop, win_length, window, center=True, onesided=True, return_complex=True)
    return torch.sqrt(torch.clamp(s.real ** 2 + s.imag ** 2, min=1e-7)).transpose(2, 1)


class MultiResolutionSTFTLoss(nn.Module):
    def __init__(self, fft_sizes=(1024, 2048, 512), hop_sizes=(120, 240, 50), win_lengths=(600, 1200, 240)):
        super().__init__()
        self.res = list(zip(fft_sizes, hop_sizes, win_lengths))
        for i, wl in enumerate(win_lengths):
            self.register_buffer(f"window{i}", torch.hann_window(wl))

    def forward(self, x, y):
        if x.dim() == 3:
            x, y = x.reshape(-1, x.size(2)), y.reshape(-1, y.size(2))
        sc, mag = 0.0, 0.0
        for i, (fs, ss, wl) in enumerate(self.res):
            w = getattr(self, f"window{i}")
            xm, ym = _stft_mag(x, fs, ss, wl, w), _stft_mag(y, fs, ss, wl, w)
            sc = sc + torch.norm(ym - xm, p="fro") / torch.norm(ym, p="fro")
            mag = mag + F.l1_loss(torch.log(ym), torch.log(xm))
        return sc / len(self.res), mag / len(self.res)


def generator_adversarial_loss(outs):
    """LSGAN generator loss averaged over the critics (adversarial_loss.py:16-57)."""
    return sum(F.mse_loss(o, o.new_ones(o.size())) for o in outs) / len(outs)


def discriminator_adversarial_loss(outs_hat, outs):
    """LSGAN critic losses (real, fake), averaged over the critics (adversarial_loss.py:60-127)."""
    real = sum(F.mse_loss(o, o.new_ones(o.size())) for o in outs) / len(outs)
    fake = sum(F.mse_loss(o, o.new_zeros(o.size())) for o in outs_hat) / len(outs_hat)
    return real, fake
