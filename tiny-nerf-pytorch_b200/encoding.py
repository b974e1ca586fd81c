"""encoding.PositionalEncoding -- drop-in for the reference's src/encoding.py:4-33."""
import torch
import torch.nn as nn

import _engine as E
import _lazy


class _PosEnc(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, num_freqs, include_input):
        dev = E.need_cuda(x)
        xc = E.f32c(x).reshape(-1, 3)
        n = xc.shape[0]
        D = 6 * num_freqs + (3 if include_input else 0)
        out = torch.empty((n, D), dtype=torch.float32, device=dev)
        E.check(E.lib().tnerf_posenc(E.ptr(xc), n, num_freqs, int(include_input), E.ptr(out), E.stream(dev)), "tnerf_posenc")
        ctx.save_for_backward(xc)
        ctx.cfg = (num_freqs, include_input, tuple(x.shape))
        return out.reshape(*x.shape[:-1], D)

    @staticmethod
    def backward(ctx, g):
        (xc,) = ctx.saved_tensors
        L, inc, shape = ctx.cfg
        gc = E.f32c(g).reshape(xc.shape[0], -1)
        gx = torch.empty_like(xc)
        E.check(E.lib().tnerf_posenc_bwd(E.ptr(xc), E.ptr(gc), xc.shape[0], L, int(inc), E.ptr(gx), E.stream(xc.device)),
                "tnerf_posenc_bwd")
        return gx.reshape(shape), None, None


def posenc_apply(x: torch.Tensor, num_freqs: int, include_input: bool) -> torch.Tensor:
    return _PosEnc.apply(x, int(num_freqs), bool(include_input))


class PositionalEncoding(nn.Module):
    """NeRF Fourier features of 3-D coordinates: ``[x, sin(2^k x), cos(2^k x)]_{k<L}``, with the column
    order of the reference (3 + 6k + 3*{sin:0, cos:1} + axis).  ``out_dim = 6L (+3)``."""

    def __init__(self, num_freqs: int = 10, include_input: bool = True):
        super().__init__()
        self.num_freqs = num_freqs
        self.include_input = include_input
        self.register_buffer("freq_bands", torch.pow(2.0, torch.arange(num_freqs).float()))

    @property
    def out_dim(self) -> int:
        return 6 * self.num_freqs + (3 if self.include_input else 0)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        assert x.shape[-1] == 3, "PositionalEncoding expects (..., 3)"
        if isinstance(x, _lazy.Deferred) and x._kind == "pts" and x._node.encoder is None and self.num_freqs <= 30:
            x._node.encoder = self
            return _lazy.Deferred((*x.shape[:-1], self.out_dim), x.device, x._node, "enc")
        return posenc_apply(x, self.num_freqs, self.include_input)
