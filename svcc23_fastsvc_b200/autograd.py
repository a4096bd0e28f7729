"""Differentiable wrapper of the CUDA generator (training callers: reference train_fastsvc.py:168,199-206).

Forward AND backward run in libfsvc.so (``fsvc_forward_train`` / ``fsvc_backward``, csrc/train.cu): the forward keeps
its activations in a buffer owned by the autograd node, the backward returns the gradient of every *effective* conv
weight.  The effective weights are inputs of the autograd function, so with weight norm applied
(``w = g * v / ||v||``, fastsvc.py:354-362) stock autograd carries the gradients on to ``weight_g`` / ``weight_v`` --
that thin elementwise chain is the only PyTorch arithmetic on the path.  Inputs (PPG, excitation, loudness, speaker
embedding) receive no gradient; the reference never asks for one.
"""

import torch

from .layers import effective_weight


def effective_weights(module, handle):
    """The tensors libfsvc expects (canonical order of ``fsvc_weight_tensor_info``), differentiable w.r.t. the
    module's parameters when grad is enabled."""
    out = []
    for name in handle.weight_names:
        path, kind = name.rsplit(".", 1)
        mod = module.get_submodule(path)
        if kind == "bias":
            out.append(mod.bias)
        elif isinstance(mod, torch.nn.Linear):
            out.append(mod.weight)
        else:
            out.append(effective_weight(mod))
    return out


class _GeneratorFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, module, x, s, l, spk, *weights):
        from .generator import _f32c, _stream
        device = x.device
        B, frames, T, spk = module._check_inputs(x, s, l, spk)
        x, s, l, spk = map(_f32c, (x, s, l, spk))
        w32 = [_f32c(w) for w in weights]
        with torch.cuda.device(device):
            handle = module._engine(device)
            for t, n, name in zip(w32, handle.weight_numel, handle.weight_names):
                if t.numel() != n:
                    raise RuntimeError(f"parameter {name} has {t.numel()} elements, library expects {n}")
            stream = _stream(device)
            handle.set_weights([t.data_ptr() for t in w32], stream)
            module._weights_key = module._param_key()
            module._weights_epoch += 1
            saved = torch.empty(handle.train_saved_bytes(B, frames), dtype=torch.uint8, device=device)
            ws = module._ws.get(handle.train_workspace_bytes(B, frames), device)
            out = torch.empty((B, module.out_channels, T), dtype=torch.float32, device=device)
            handle.forward_train(x.data_ptr(), s.data_ptr(), l.data_ptr(), 0 if spk is None else spk.data_ptr(),
                                 out.data_ptr(), B, frames, saved.data_ptr(), saved.numel(), ws.data_ptr(),
                                 ws.numel(), stream)
        ctx.module, ctx.handle, ctx.epoch = module, handle, module._weights_epoch
        ctx.dims = (B, frames)
        ctx.saved_buf = saved
        ctx.inputs = (x, s, l, spk)
        ctx.wmeta = [(w.shape, w.dtype) for w in weights]
        return out

    @staticmethod
    def backward(ctx, grad_out):
        from .generator import _f32c, _stream
        module, handle = ctx.module, ctx.handle
        if module._weights_epoch != ctx.epoch or module._handle is not handle:
            raise RuntimeError(
                "FastSVCGenerator: the weights inside libfsvc changed between this forward and its backward (another "
                "grad-enabled forward, an optimizer step followed by a forward, or a device move); gradients would "
                "belong to different weights. Run backward before the next forward.")
        x, s, l, spk = ctx.inputs
        B, frames = ctx.dims
        device = x.device
        g = _f32c(grad_out)
        grads = [torch.empty(shape, dtype=torch.float32, device=device) for shape, _ in ctx.wmeta]
        with torch.cuda.device(device):
            ws = module._ws.get(handle.train_workspace_bytes(B, frames), device)
            handle.backward(x.data_ptr(), s.data_ptr(), l.data_ptr(), 0 if spk is None else spk.data_ptr(),
                            g.data_ptr(), B, frames, ctx.saved_buf.data_ptr(), ctx.saved_buf.numel(),
                            [t.data_ptr() for t in grads], ws.data_ptr(), ws.numel(), _stream(device))
        ctx.saved_buf = None
        grads = [gr if dt == torch.float32 else gr.to(dt) for gr, (_, dt) in zip(grads, ctx.wmeta)]
        if spk is None:  # emb_projector takes no part in the graph (fastsvc.py:134): no gradient, like stock autograd
            grads = [None if ".emb_projector." in name else gr for gr, name in zip(grads, handle.weight_names)]
        return (None, None, None, None, None, *grads)


def generator_forward_with_grad(module, x, s, l, spk):
    for name, t in (("x", x), ("s", s), ("l", l), ("spk_emb", spk)):
        if t is not None and t.requires_grad:
            raise NotImplementedError(
                f"FastSVCGenerator (B200-native): gradient w.r.t. the input {name!r} is not implemented; the native "
                "backward returns parameter gradients only (all the reference's training loop uses)")
    handle = module._engine(x.device)
    return _GeneratorFn.apply(module, x, s, l, spk, *effective_weights(module, handle))
