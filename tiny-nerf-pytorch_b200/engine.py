"""Public fused API of the B200 engine: whole-ray rendering / training-step calls that map 1:1 onto
the fused C-ABI entry points (tnerf_render_fwd / tnerf_render_bwd / tnerf_train_fwd_bwd /
tnerf_adam_step).  volume.volume_render reaches the same code through deferred tensors, so the
reference's scripts use it without knowing."""
from __future__ import annotations

import ctypes as C
import math
import os
from typing import Optional

import torch

import _engine as E

_PREC = {"f16": E.PREC_F16_TC, "fp16": E.PREC_F16_TC, "tc": E.PREC_F16_TC, "f32": E.PREC_F32_SIMT, "fp32": E.PREC_F32_SIMT}


def default_precision() -> int:
    return _PREC[os.environ.get("TNERF_PRECISION", "f16").lower()]


def default_bwd_precision() -> int:
    return _PREC[os.environ.get("TNERF_BWD_PRECISION", "f32").lower()]


def ray_source(rays_o=None, o_stride=3, rays_d=None, c2w=None, H=0, W=0, focal=0.0, pixel_index=None, first_ray=0):
    rs = E.RaySource()
    rs.rays_o = E.ptr(rays_o); rs.o_stride = int(o_stride); rs.rays_d = E.ptr(rays_d); rs.c2w = E.ptr(c2w)
    rs.H, rs.W, rs.focal = int(H), int(W), float(focal)
    rs.pixel_index = E.ptr(pixel_index); rs.first_ray = int(first_ray)
    return rs


def origin_arg(rays_o):
    """(tensor, row stride) for the C ABI: a broadcast origin (get_rays' expand view, stride 0 over rays)
    is passed as its single row, anything else as a dense (N,3) fp32 tensor."""
    if rays_o.dim() == 2 and rays_o.shape[0] > 0 and rays_o.stride(0) == 0:
        return E.f32c(rays_o[0]), 0
    return E.f32c(rays_o), 3


class _FusedRender(torch.autograd.Function):
    """comp_rgb, depth, acc = fused(rays, samples, MLP).  Backward recomputes activations on chip."""

    @staticmethod
    def forward(ctx, module, ro, o_stride, rd, n, S, near, far, jitter, white, prec, bwd_prec, *params):
        dev = rd.device
        h = E.handle_for(module, dev)
        if prec == E.PREC_F16_TC:
            h.ensure_packed()
        else:
            h.bind()
        comp = torch.empty((n, 3), dtype=torch.float32, device=dev)
        depth = torch.empty((n, 1), dtype=torch.float32, device=dev)
        acc = torch.empty((n, 1), dtype=torch.float32, device=dev)
        rs = ray_source(ro, o_stride, rd)
        E.check(E.lib().tnerf_render_fwd(h.h, C.byref(rs), n, near, far, S, E.ptr(jitter), int(white), prec, E.ptr(comp),
                                         E.ptr(depth), E.ptr(acc), None, None, E.stream(dev)), "tnerf_render_fwd")
        ctx.args = (module, ro, o_stride, rd, n, S, near, far, jitter, white, bwd_prec)
        return comp, depth, acc

    @staticmethod
    def backward(ctx, gC, gD, gA):
        module, ro, o_stride, rd, n, S, near, far, jitter, white, prec = ctx.args
        dev = rd.device
        h = E.handle_for(module, dev)
        ps = h.bind()
        if prec == E.PREC_F16_TC:
            h.ensure_packed()
        grads = torch.zeros(h.param_count, dtype=torch.float32, device=dev)
        rs = ray_source(ro, o_stride, rd)
        gC, gD, gA = (E.f32c(g) if g is not None else None for g in (gC, gD, gA))
        E.check(E.lib().tnerf_render_bwd(h.h, C.byref(rs), n, near, far, S, E.ptr(jitter), int(white), prec, E.ptr(gC), E.ptr(gD),
                                         E.ptr(gA), None, 0.0, E.ptr(grads), E.stream(dev)), "tnerf_render_bwd")
        views = E.flat_grad_views(module, grads)
        return (None,) * 12 + tuple(v if p.requires_grad else None for v, p in zip(views, ps))


def fused_supported(module, encoder, S: int, device) -> bool:
    h = E.handle_for(module, device)
    if encoder is not None and encoder.out_dim == module.in_dim:
        try:
            h.set_encoding(encoder.num_freqs, encoder.include_input)
        except RuntimeError:
            return False
    g = math.gcd(int(S), 128)
    return h.fused_ok and module.hidden == 128 and S // g <= 8


def render_rays(model, encoder, rays_o, rays_d, near: float, far: float, n_samples: int, t_rand: Optional[torch.Tensor] = None,
                white_bkgd: bool = True, precision: Optional[str] = None):
    """Fused a3->a6: returns (comp_rgb (N,3), depth (N,1), acc (N,1)).  Differentiable w.r.t. the MLP
    parameters.  ``t_rand`` (N,S) enables stratified jitter; None = deterministic depths."""
    dev = E.need_cuda(rays_o, rays_d)
    prec = _PREC[precision.lower()] if precision else default_precision()
    bprec = _PREC[precision.lower()] if precision else default_bwd_precision()
    h = E.handle_for(model, dev)
    h.set_encoding(encoder.num_freqs, encoder.include_input)
    if prec == E.PREC_F16_TC and not fused_supported(model, encoder, n_samples, dev):
        prec = bprec = E.PREC_F32_SIMT
    ro, o_stride = origin_arg(rays_o)
    rd = E.f32c(rays_d)
    jit = E.f32c(t_rand) if t_rand is not None else None
    return _FusedRender.apply(model, ro, o_stride, rd, int(rd.shape[0]), int(n_samples), float(near), float(far), jit,
                              bool(white_bkgd), prec, bprec, *model._params())


def render_weights(model, ro, o_stride, rd, n, S, near, far, jitter, white, prec):
    """The (N,S) compositing weights of a fused render (the 4th value of volume_render), recomputed on demand."""
    dev = rd.device
    h = E.handle_for(model, dev)
    if prec == E.PREC_F16_TC:
        h.ensure_packed()
    comp = torch.empty((n, 3), dtype=torch.float32, device=dev)
    w = torch.empty((n, S), dtype=torch.float32, device=dev)
    rs = ray_source(ro, o_stride, rd)
    E.check(E.lib().tnerf_render_fwd(h.h, C.byref(rs), n, near, far, S, E.ptr(jitter), int(white), prec, E.ptr(comp), None, None,
                                     E.ptr(w), None, E.stream(dev)), "tnerf_render_fwd")
    return w
