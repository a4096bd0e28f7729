"""Drop-in ``FastSVCGenerator`` and its sub-blocks, computed by libfsvc.so.

Mirrors the reference's class / constructor / ``forward`` / ``inference`` /
weight-norm / ``state_dict`` surface (harana/models/fastsvc.py:34-383) so that
``getattr(harana.models, "FastSVCGenerator")(**generator_params)``
(train_fastsvc.py:700-713, utils/utils.py:266-275) and reference checkpoints
work unchanged.  The module tree only *holds parameters*; all arithmetic of
``forward`` happens in hand-written sm_100a CUDA kernels behind the C ABI of
``include/fsvc.h``.  There is no CPU or PyTorch-op fallback: CPU tensors, a
missing library or a missing GPU raise.
"""

import os

import torch
from torch import nn

from . import abi
from .layers import Conv1d1x1, Conv1d1x3, Conv2d1x3, Squeeze2d, Stretch2d, effective_weight

LRELU_SLOPE = 0.2
IN_EPS = 1e-5


def _require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError(
                "FastSVC (B200-native) runs only on CUDA tensors: there is no CPU fallback. "
                "Move the module and its inputs to a CUDA device.")


def _f32c(t):
    return None if t is None else t.detach().to(torch.float32).contiguous()


def _stream(device):
    return torch.cuda.current_stream(device).cuda_stream


class _Workspace:
    """Grow-only device scratch buffer owned by a module (not part of state_dict)."""

    def __init__(self):
        self.buf = None

    def get(self, nbytes, device):
        if self.buf is None or self.buf.numel() < nbytes or self.buf.device != device:
            self.buf = None
            self.buf = torch.empty(int(nbytes), dtype=torch.uint8, device=device)
        return self.buf


def _block_mode(module):
    """Precision mode of a standalone sub-block: attribute ``precision`` or $FSVC_MODE, default auto."""
    name = getattr(module, "precision", None) or os.environ.get("FSVC_MODE", "auto")
    try:
        return abi.MODES[name]
    except KeyError:
        raise ValueError(f"precision must be one of {sorted(abi.MODES)}, got {name!r}")


def _conv_wb(conv):
    return [_f32c(effective_weight(conv)), _f32c(conv.bias)]


def _refuse_grad(module, *tensors):
    """The standalone sub-blocks are inference-only entry points (their outputs carry no ``grad_fn``): refuse a
    call that expects gradients instead of dropping them silently.  Training goes through ``FastSVCGenerator``."""
    if torch.is_grad_enabled() and (any(p.requires_grad for p in module.parameters())
                                    or any(t is not None and t.requires_grad for t in tensors)):
        raise RuntimeError(
            f"{type(module).__name__}.forward is inference-only in the B200-native build (no autograd through the "
            "standalone block); call it under torch.no_grad(), or train through FastSVCGenerator.forward, which has "
            "a native backward.")


class FastSVCUpsampleNet(nn.Module):
    """FastSVC upsampling block (reference fastsvc.py:34-140).

    conv_first -> {residual: repeat+conv ; main: lrelu, repeat, conv, lrelu} ->
    FiLM/InstanceNorm/speaker add -> dilated convs 3, 9, 27 with two skips.
    """

    def __init__(self, in_channels, mid_channels, upsampling_scale, spk_emb_size=512, use_spk_emb=True):
        super().__init__()
        self.conv_first = Conv2d1x3(in_channels, mid_channels, (0, 1), 1)
        self.upsample_block0 = nn.Sequential(
            nn.LeakyReLU(LRELU_SLOPE), Stretch2d(upsampling_scale, 1),
            Conv2d1x3(mid_channels, mid_channels, (0, 1), 1), nn.LeakyReLU(LRELU_SLOPE))
        self.conv_block1 = nn.Sequential(nn.LeakyReLU(LRELU_SLOPE), Conv2d1x3(mid_channels, mid_channels, (0, 3), 3))
        self.conv_block2 = nn.Sequential(nn.LeakyReLU(LRELU_SLOPE), Conv2d1x3(mid_channels, mid_channels, (0, 9), 9))
        self.conv_block3 = nn.Sequential(nn.LeakyReLU(LRELU_SLOPE),
                                         Conv2d1x3(mid_channels, mid_channels, (0, 27), 27))
        self.residual_block = nn.Sequential(Stretch2d(upsampling_scale, 1),
                                            Conv2d1x3(mid_channels, mid_channels, (0, 1), 1))
        self.instance_norm = nn.InstanceNorm2d(mid_channels)
        if use_spk_emb:
            self.emb_projector = nn.Linear(spk_emb_size, mid_channels)
        self._scale = upsampling_scale
        self._ws = _Workspace()

    def forward(self, x, s, l, spk_emb=None):
        """x (B, C_in, T); s = (scale, shift), l = (scale, shift), each (B, C, T*r); -> (B, C, T*r)."""
        s_scale, s_shift = s
        l_scale, l_shift = l
        _refuse_grad(self, x, s_scale, s_shift, l_scale, l_shift, spk_emb)
        _require_cuda(x, s_scale, s_shift, l_scale, l_shift, spk_emb)
        lib = abi.load()
        x, s_scale, s_shift, l_scale, l_shift, spk_emb = map(_f32c, (x, s_scale, s_shift, l_scale, l_shift, spk_emb))
        B, c_in, T = x.shape
        c = self.conv_first.out_channels
        w = []
        for conv in (self.conv_first, self.upsample_block0[2], self.conv_block1[1], self.conv_block2[1],
                     self.conv_block3[1], self.residual_block[1]):
            w += _conv_wb(conv)
        spk_size = 0
        if spk_emb is not None:
            if not hasattr(self, "emb_projector"):
                raise ValueError("spk_emb given but the block was built with use_spk_emb=False")
            w += [_f32c(self.emb_projector.weight), _f32c(self.emb_projector.bias)]
            spk_size = self.emb_projector.in_features
        ptrs = [t.data_ptr() for t in w] + [0] * (14 - len(w))
        out = torch.empty((B, c, T * self._scale), dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            nbytes = lib.fsvc_block_workspace_bytes(B, c_in, c, T * self._scale)
            ws = self._ws.get(nbytes, x.device)
            abi.check(lib.fsvc_upsample_forward(
                x.data_ptr(), s_scale.data_ptr(), s_shift.data_ptr(), l_scale.data_ptr(), l_shift.data_ptr(),
                0 if spk_emb is None else spk_emb.data_ptr(), out.data_ptr(), abi.ptr_array(ptrs), B, c_in, c, T,
                self._scale, spk_size, LRELU_SLOPE, IN_EPS, ws.data_ptr(), ws.numel(), _block_mode(self),
                _stream(x.device)))
        return out


class FastSVCDownsampleNet(nn.Module):
    """FastSVC downsampling block (reference fastsvc.py:143-193):
    decimate, three dilated (1, 2, 4) k=3 convs, plus a decimated 1x1 residual."""

    def __init__(self, in_channels=1, mid_channels=[12, 24, 48, 96, 192], downsampling_scales=[1, 5, 4, 4, 4]):
        super().__init__()
        self.residual_block = nn.Sequential(Conv1d1x1(in_channels, mid_channels), Squeeze2d(downsampling_scales))
        self.downsample_block = nn.Sequential(
            Squeeze2d(downsampling_scales), nn.LeakyReLU(LRELU_SLOPE), Conv1d1x3(in_channels, mid_channels, 1, 1),
            nn.LeakyReLU(LRELU_SLOPE), Conv1d1x3(mid_channels, mid_channels, 2, 2),
            nn.LeakyReLU(LRELU_SLOPE), Conv1d1x3(mid_channels, mid_channels, 4, 4))
        self._scale = downsampling_scales
        self._ws = _Workspace()

    def forward(self, x):
        """x (B, C_in, T) -> (B, C, T / scale); T must be divisible by the scale."""
        _refuse_grad(self, x)
        _require_cuda(x)
        lib = abi.load()
        x = _f32c(x)
        B, c_in, T = x.shape
        c = self.residual_block[0].out_channels
        if T % self._scale:
            raise ValueError(f"T={T} must be divisible by the downsampling scale {self._scale}")
        w = []
        for conv in (self.residual_block[0], self.downsample_block[2], self.downsample_block[4],
                     self.downsample_block[6]):
            w += _conv_wb(conv)
        out = torch.empty((B, c, T // self._scale), dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            ws = self._ws.get(lib.fsvc_block_workspace_bytes(B, c_in, c, T // self._scale), x.device)
            abi.check(lib.fsvc_downsample_forward(
                x.data_ptr(), out.data_ptr(), abi.ptr_array([t.data_ptr() for t in w]), B, c_in, c, T, self._scale,
                LRELU_SLOPE, ws.data_ptr(), ws.numel(), _block_mode(self), _stream(x.device)))
        return out


class FastSVCFiLMNet(nn.Module):
    """FastSVC FiLM block (reference fastsvc.py:196-232): conv, lrelu, then a scale conv and a shift conv."""

    def __init__(self, mid_channels):
        super().__init__()
        self.conv = Conv1d1x3(mid_channels, mid_channels, padding=1, dilation=1)
        self.relu = nn.LeakyReLU(LRELU_SLOPE)
        self.conv_scale = Conv1d1x3(mid_channels, mid_channels, padding=1, dilation=1)
        self.conv_shift = Conv1d1x3(mid_channels, mid_channels, padding=1, dilation=1)
        self._ws = _Workspace()

    def forward(self, x):
        """x (B, C, T) -> (scale, shift), each (B, C, T)."""
        _refuse_grad(self, x)
        _require_cuda(x)
        lib = abi.load()
        x = _f32c(x)
        B, c, T = x.shape
        w = _conv_wb(self.conv) + _conv_wb(self.conv_scale) + _conv_wb(self.conv_shift)
        scale, shift = torch.empty_like(x), torch.empty_like(x)
        with torch.cuda.device(x.device):
            ws = self._ws.get(lib.fsvc_block_workspace_bytes(B, c, c, T), x.device)
            abi.check(lib.fsvc_film_forward(
                x.data_ptr(), scale.data_ptr(), shift.data_ptr(), abi.ptr_array([t.data_ptr() for t in w]), B, c, T,
                LRELU_SLOPE, ws.data_ptr(), ws.numel(), _block_mode(self), _stream(x.device)))
        return scale, shift


class FastSVCGenerator(nn.Module):
    """FastSVC waveform generator (reference fastsvc.py:235-383).

    PPG (B, C_in, T') + sine (B, 1, T) + loudness (B, 1, T) [+ speaker embedding
    (B, S)] -> waveform (B, out_channels, T), T = T' * prod(upsampling_scales).
    """

    def __init__(self, in_channels=144, mid_channels=[192, 96, 48, 24], upsampling_scales=[2, 4, 4, 5],
                 out_channels=1, spk_emb_size=512, use_spk_emb=True):
        super().__init__()
        if len(mid_channels) != len(upsampling_scales):
            raise ValueError("mid_channels and upsampling_scales must have the same length")
        self.in_channels = in_channels
        self.upsampling_scales = upsampling_scales
        self.mid_channels = mid_channels
        self.out_channels = out_channels
        self.spk_emb_size = spk_emb_size
        self.use_spk_emb = use_spk_emb

        self.upsampling_nets = nn.ModuleList()
        cin = in_channels
        for scale, channel in zip(upsampling_scales, mid_channels):
            self.upsampling_nets.append(FastSVCUpsampleNet(cin, channel, scale, spk_emb_size, use_spk_emb))
            cin = channel

        # conditioning chains run coarse-to-fine in reverse stage order (fastsvc.py:270-287)
        down_scales = [1] + list(upsampling_scales)[::-1][:-1]
        lft_layers, sine_layers = [], []
        cin = 1
        for scale, channel in zip(down_scales, list(mid_channels)[::-1]):
            lft_layers.append(FastSVCDownsampleNet(cin, channel, scale))
            sine_layers.append(FastSVCDownsampleNet(cin, channel, scale))
            cin = channel
        self.downsampling_lft = nn.Sequential(*lft_layers)
        self.downsampling_sine = nn.Sequential(*sine_layers)

        self.film_lft = nn.ModuleList()
        self.film_sine = nn.ModuleList()
        for channel in list(mid_channels)[::-1]:
            self.film_lft.append(FastSVCFiLMNet(channel))
            self.film_sine.append(FastSVCFiLMNet(channel))

        self.conv_last = Conv1d1x1(mid_channels[-1], out_channels)
        self.apply_weight_norm()

        # runtime state (never in state_dict)
        self.precision = os.environ.get("FSVC_MODE", "auto")
        self._handle = None
        self._handle_device = None
        self._weights_key = None
        self._weights_epoch = 0     # bumped whenever new weights are pushed into the library (autograd.py checks it)
        self._ws = _Workspace()

    # ---- weight norm (fastsvc.py:342-362) ----
    def remove_weight_norm(self):
        def _remove(m):
            try:
                torch.nn.utils.remove_weight_norm(m)
            except ValueError:
                return

        self.apply(_remove)

    def apply_weight_norm(self):
        def _apply(m):
            if isinstance(m, (nn.Conv1d, nn.Conv2d)) and not hasattr(m, "weight_g"):
                torch.nn.utils.weight_norm(m)

        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore", FutureWarning)
            self.apply(_apply)

    # ---- engine ----
    @property
    def hop_size(self):
        h = 1
        for r in self.upsampling_scales:
            h *= r
        return h

    def _engine(self, device):
        if self._handle is None or self._handle_device != device:
            if self._handle is not None:
                self._handle.close()
            with torch.cuda.device(device):
                self._handle = abi.Handle(self.in_channels, list(self.mid_channels), list(self.upsampling_scales),
                                          self.out_channels, self.spk_emb_size, self.use_spk_emb, LRELU_SLOPE, IN_EPS)
            self._handle_device = device
            self._weights_key = None
        return self._handle

    def _sync_weights(self, handle, device):
        """Push effective weights to the library when any parameter changed
        (optimizer step, load_state_dict, remove/apply_weight_norm, .to())."""
        key = self._param_key()
        if key == self._weights_key:
            return
        from .autograd import effective_weights
        tensors = []
        with torch.no_grad():
            for name, t in zip(handle.weight_names, effective_weights(self, handle)):
                if t.device != device:
                    raise RuntimeError(f"parameter {name} is on {t.device}, inputs on {device}")
                tensors.append(_f32c(t))
        for t, n, name in zip(tensors, handle.weight_numel, handle.weight_names):
            if t.numel() != n:
                raise RuntimeError(f"parameter {name} has {t.numel()} elements, library expects {n}")
        handle.set_weights([t.data_ptr() for t in tensors], _stream(device))
        self._weights_key = key
        self._weights_epoch += 1

    def _param_key(self):
        return tuple((p.data_ptr(), p._version) for p in self.parameters())

    def _mode(self):
        try:
            return abi.MODES[self.precision]
        except KeyError:
            raise ValueError(f"precision must be one of {sorted(abi.MODES)}, got {self.precision!r}")

    def forward(self, x, s, l, spk_emb=None):
        """x (B, C_in, T'), s (B, 1, T), l (B, 1, T), spk_emb (B, S) or None -> (B, out_channels, T).
        ``spk_emb=None`` applies the FiLM affine without InstanceNorm / speaker add (fastsvc.py:134)."""
        _require_cuda(x, s, l, spk_emb)
        if torch.is_grad_enabled() and (any(p.requires_grad for p in self.parameters())
                                        or x.requires_grad or s.requires_grad or l.requires_grad):
            from .autograd import generator_forward_with_grad
            return generator_forward_with_grad(self, x, s, l, spk_emb)
        return self._forward_cuda(x, s, l, spk_emb)

    def _check_inputs(self, x, s, l, spk_emb):
        """Shape validation shared by every entry point; returns (B, frames, T, spk_emb expanded to B rows)."""
        if x.dim() != 3 or s.dim() != 3 or l.dim() != 3:
            raise ValueError("x, s, l must be (B, C, T) tensors")
        B, cin, frames = x.shape
        T = frames * self.hop_size
        if cin != self.in_channels:
            raise ValueError(f"x has {cin} channels, generator expects {self.in_channels}")
        if tuple(s.shape) != (B, 1, T) or tuple(l.shape) != (B, 1, T):
            raise ValueError(f"s and l must be (B, 1, T' * {self.hop_size}) = ({B}, 1, {T}); got "
                             f"{tuple(s.shape)} and {tuple(l.shape)}")
        if spk_emb is not None:
            if not self.use_spk_emb:
                raise ValueError("spk_emb given but the generator was built with use_spk_emb=False")
            if spk_emb.dim() != 2 or spk_emb.shape[1] != self.spk_emb_size or spk_emb.shape[0] not in (1, B):
                raise ValueError(f"spk_emb must be ({B}, {self.spk_emb_size}), got {tuple(spk_emb.shape)}")
            if spk_emb.shape[0] != B:
                spk_emb = spk_emb.expand(B, -1)   # the (1, S) target speaker of decode_fastsvc.py:156-158
        return B, frames, T, spk_emb

    def _forward_cuda(self, x, s, l, spk_emb=None):
        B, frames, T, spk_emb = self._check_inputs(x, s, l, spk_emb)
        device = x.device
        x, s, l, spk_emb = map(_f32c, (x, s, l, spk_emb))
        with torch.cuda.device(device):
            handle = self._engine(device)
            self._sync_weights(handle, device)
            mode = self._mode()
            nbytes = handle.workspace_bytes(B, frames, mode)
            ws = self._ws.get(nbytes, device)
            out = torch.empty((B, self.out_channels, T), dtype=torch.float32, device=device)
            handle.forward(x.data_ptr(), s.data_ptr(), l.data_ptr(), 0 if spk_emb is None else spk_emb.data_ptr(),
                           out.data_ptr(), B, frames, ws.data_ptr(), ws.numel(), mode, _stream(device))
        return out

    def forward_host(self, x, s, l, spk_emb=None, out=None):
        """End-to-end call on HOST tensors (pinned for asynchrony): H2D copies,
        forward, D2H copy of the waveform, all enqueued on the current stream of
        the module's device by ``fsvc_forward_host``.  Returns the (pinned) host
        output tensor; synchronise the stream before reading it."""
        for t in (x, s, l, spk_emb):
            if t is not None and t.is_cuda:
                raise ValueError("forward_host takes host tensors")
        B, frames, T, spk_emb = self._check_inputs(x, s, l, spk_emb)
        device = next(self.parameters()).device
        if device.type != "cuda":
            raise RuntimeError("FastSVC (B200-native) needs its parameters on a CUDA device: no CPU fallback")
        x, s, l, spk_emb = map(_f32c, (x, s, l, spk_emb))   # .contiguous() materialises an expanded (1, S) speaker
        if out is None:
            out = torch.empty((B, self.out_channels, T), dtype=torch.float32).pin_memory()
        elif tuple(out.shape) != (B, self.out_channels, T) or out.dtype != torch.float32 or out.is_cuda \
                or not out.is_contiguous():
            raise ValueError(f"out must be a contiguous host fp32 tensor of shape {(B, self.out_channels, T)}")
        with torch.cuda.device(device):
            handle = self._engine(device)
            self._sync_weights(handle, device)
            mode = self._mode()
            nbytes = handle.workspace_bytes(B, frames, mode) + handle.host_io_bytes(B, frames)
            ws = self._ws.get(nbytes, device)
            handle.forward_host(x.data_ptr(), s.data_ptr(), l.data_ptr(),
                                0 if spk_emb is None else spk_emb.data_ptr(), out.data_ptr(), B, frames,
                                ws.data_ptr(), ws.numel(), mode, _stream(device))
        return out

    def profile(self, x, s, l, spk_emb=None):
        """Per-launch device times of one forward (``fsvc_forward_profile``): list of dicts
        {label, ms, flops, bytes}.  Synchronises; for benchmarking only."""
        _require_cuda(x, s, l, spk_emb)
        B, frames, _, spk_emb = self._check_inputs(x, s, l, spk_emb)
        device = x.device
        x, s, l, spk_emb = map(_f32c, (x, s, l, spk_emb))
        with torch.cuda.device(device):
            handle = self._engine(device)
            self._sync_weights(handle, device)
            mode = self._mode()
            ws = self._ws.get(handle.workspace_bytes(B, frames, mode), device)
            out = torch.empty((B, self.out_channels, frames * self.hop_size), dtype=torch.float32, device=device)
            return handle.forward_profile(x.data_ptr(), s.data_ptr(), l.data_ptr(),
                                          0 if spk_emb is None else spk_emb.data_ptr(), out.data_ptr(), B, frames,
                                          ws.data_ptr(), ws.numel(), mode, _stream(device))

    def last_launch_count(self):
        return 0 if self._handle is None else self._handle.last_launch_count()

    def __getstate__(self):
        # the library handle / scratch are per-process runtime state: never copied or pickled
        state = self.__dict__.copy()
        state["_handle"] = None
        state["_handle_device"] = None
        state["_weights_key"] = None
        state["_weights_epoch"] = 0
        state["_ws"] = _Workspace()
        return state

    def downsampling_loop(self, x, didx, nets):
        """Output of the conditioning chain after level ``didx`` (fastsvc.py:334-340)."""
        for idx, net in enumerate(nets):
            x = net(x)
            if idx == didx:
                return x
        raise ValueError("index went over the length of the network")

    def inference(self, x, f0, l, signal_generator, pad_fn, spk_emb=None):
        """Single-utterance inference (fastsvc.py:364-383): x (T', C), f0 (T', 1), l (T, 1) -> (T, out_channels)."""
        x = pad_fn(x.transpose(1, 0).unsqueeze(0))
        l = l.transpose(1, 0).unsqueeze(0)
        f0 = f0.transpose(1, 0).unsqueeze(0)
        s = signal_generator(f0)
        return self.forward(x, s, l, spk_emb).squeeze(0).transpose(1, 0)
