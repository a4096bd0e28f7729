"""Differentiable wrapper of the CUDA generator forward (training callers:
reference train_fastsvc.py:168,199-206).

Forward values come from libfsvc.so.  INTERIM backward (SURVEY.md 8f row N1 --
the native CUDA backward is the next row, not part of the forward hot path):
the graph is rebuilt in ``backward`` from stock PyTorch CUDA ops over the
module's live parameters and differentiated with ``torch.autograd.grad``.  It
runs on the GPU only and never produces forward values.
"""

import torch
import torch.nn.functional as F

from .layers import effective_weight

_SLOPE = 0.2


def _c(conv, x, dil):
    w = effective_weight(conv)
    if w.dim() == 4:
        w = w[:, :, 0, :]
    k = w.shape[-1]
    return F.conv1d(x, w, conv.bias, dilation=dil, padding=dil * (k // 2))


def _down(net, x, scale):
    xd = x[..., ::scale]
    r = _c(net.residual_block[0], xd, 1)
    h = _c(net.downsample_block[2], F.leaky_relu(xd, _SLOPE), 1)
    h = _c(net.downsample_block[4], F.leaky_relu(h, _SLOPE), 2)
    h = _c(net.downsample_block[6], F.leaky_relu(h, _SLOPE), 4)
    return h + r


def _film(net, y):
    h = F.leaky_relu(_c(net.conv, y, 1), _SLOPE)
    return _c(net.conv_scale, h, 1), _c(net.conv_shift, h, 1)


def _stage(net, x, gamma, beta, r, spk):
    def fa(t):
        t = gamma * t + beta
        if spk is not None:
            t = F.instance_norm(t, eps=1e-5) + net.emb_projector(F.normalize(spk)).unsqueeze(-1)
        return F.leaky_relu(t, _SLOPE)

    h0 = _c(net.conv_first, x, 1)
    xr = _c(net.residual_block[1], h0.repeat_interleave(r, -1), 1)
    u = F.leaky_relu(_c(net.upsample_block0[2], F.leaky_relu(h0, _SLOPE).repeat_interleave(r, -1), 1), _SLOPE)
    x_ = _c(net.conv_block1[1], fa(u), 3) + xr
    x2 = _c(net.conv_block2[1], fa(x_), 9)
    return _c(net.conv_block3[1], fa(x2), 27) + x_


def _graph_forward(g, x, s, l, spk):
    n = len(g.upsampling_nets)
    scales = list(g.upsampling_scales)
    down = [1] + scales[::-1][:-1]
    gb = []
    hl, hs = l, s
    for i in range(n):
        hl = _down(g.downsampling_lft[i], hl, down[i])
        hs = _down(g.downsampling_sine[i], hs, down[i])
        gl, bl = _film(g.film_lft[i], hl)
        gs, bs = _film(g.film_sine[i], hs)
        gb.append((gl + gs, bl + bs))
    for i in range(n):
        gamma, beta = gb[n - 1 - i]
        x = _stage(g.upsampling_nets[i], x, gamma, beta, scales[i], spk)
    return _c(g.conv_last, x, 1)


class _GeneratorFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, module, x, s, l, spk, *params):
        ctx.module = module
        ctx.has_spk = spk is not None
        ctx.save_for_backward(x, s, l, *([spk] if spk is not None else []))
        return module._forward_cuda(x, s, l, spk)

    @staticmethod
    def backward(ctx, grad_out):
        g = ctx.module
        saved = ctx.saved_tensors
        x, s, l = saved[:3]
        spk = saved[3] if ctx.has_spk else None
        params = [p for p in g.parameters()]
        # full fp32 in the interim graph: with cuDNN's default TF32 convolutions the parameter gradients drift ~2 %
        # from the reference's fp32 CPU autograd (tests/test_grads.py)
        with torch.enable_grad(), torch.backends.cudnn.flags(enabled=True, benchmark=torch.backends.cudnn.benchmark,
                                                             deterministic=torch.backends.cudnn.deterministic,
                                                             allow_tf32=False):
            ins = [t.detach().requires_grad_(need) for t, need in zip((x, s, l), ctx.needs_input_grad[1:4])]
            spk_in = None
            if spk is not None:
                spk_in = spk.detach().requires_grad_(ctx.needs_input_grad[4])
            y = _graph_forward(g, ins[0], ins[1], ins[2], spk_in)
            wanted = [t for t in ins + ([spk_in] if spk_in is not None else []) + params if t.requires_grad]
            grads = torch.autograd.grad(y, wanted, grad_out, allow_unused=True)
        it = iter(grads)
        res = [None]
        for t in ins:
            res.append(next(it) if t.requires_grad else None)
        res.append(next(it) if (spk_in is not None and spk_in.requires_grad) else None)
        for p in params:
            res.append(next(it) if p.requires_grad else None)
        return tuple(res)


def generator_forward_with_grad(module, x, s, l, spk):
    return _GeneratorFn.apply(module, x, s, l, spk, *module.parameters())
