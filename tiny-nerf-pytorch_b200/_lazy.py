"""Deferred tensors: how the reference's five separate calls collapse into one fused kernel.

The reference's callers (src/train.py:114-121 and :51-56, src/main.py:25-31) run
``stratified_samples -> encoder(pts.reshape(-1,3)) -> model(xenc) -> reshape -> volume_render`` as
separate ops.  To keep those scripts unchanged and still never write points / encodings /
activations to HBM, ``stratified_samples`` returns ``pts`` as a ``Deferred`` tensor: a wrapper with
the right shape/dtype/device but no storage, which remembers how it would be computed.
``PositionalEncoding`` and ``TinyNeRF`` propagate it, and ``volume_render`` recognises the complete
chain and launches the fused kernel.  ANY other use (indexing, arithmetic, printing, a different
z_vals, ...) materialises the value with the stand-alone kernels, so semantics are unchanged.
"""
from __future__ import annotations

import os
from dataclasses import dataclass
from typing import Optional

import torch
from torch.utils._pytree import tree_map

import _engine as E

FUSE = os.environ.get("TNERF_FUSE", "1") != "0"


@dataclass
class SampleSpec:
    ro: torch.Tensor
    o_stride: int
    rd: torch.Tensor
    n: int
    S: int
    near: float
    far: float
    near_t: Optional[torch.Tensor]
    far_t: Optional[torch.Tensor]
    jitter: Optional[torch.Tensor]
    z_vals: torch.Tensor


class Node:
    """One deferred sample set flowing through encoder and model."""

    def __init__(self, spec: SampleSpec):
        self.spec = spec
        self.encoder = None        # PositionalEncoding module once applied
        self.model = None          # TinyNeRF module once applied
        self.cache = {}

    # ---- eager evaluation with the stand-alone kernels ------------------------------------
    def value(self, kind: str) -> torch.Tensor:
        if kind in self.cache:
            return self.cache[kind]
        s = self.spec
        dev = s.rd.device
        if kind == "pts":
            pts = torch.empty((s.n, s.S, 3), dtype=torch.float32, device=dev)
            E.check(E.lib().tnerf_stratified(E.ptr(s.ro), s.o_stride, E.ptr(s.rd), s.n, s.S, s.near, s.far,
                                             E.ptr(s.near_t), E.ptr(s.far_t), E.ptr(s.jitter), None, E.ptr(pts),
                                             E.stream(dev)), "tnerf_stratified")
            out = pts
        elif kind == "enc":
            import encoding
            out = encoding.posenc_apply(self.value("pts").reshape(-1, 3), self.encoder.num_freqs,
                                        self.encoder.include_input)
        elif kind in ("rgb", "sigma"):
            rgb, sigma = self.model._forward_dense(self.value("enc"))
            self.cache["rgb"], self.cache["sigma"] = rgb, sigma
            return self.cache[kind]
        else:
            raise KeyError(kind)
        self.cache[kind] = out
        return out


_META = {"dim", "size", "numel", "__len__", "ndimension", "nelement", "is_contiguous", "stride", "is_floating_point",
         "is_complex", "element_size", "get_device", "data_ptr_disabled"}


class Deferred(torch.Tensor):
    @staticmethod
    def __new__(cls, shape, device, node: Node, kind: str):
        t = torch.Tensor._make_wrapper_subclass(cls, tuple(int(v) for v in shape), dtype=torch.float32,
                                                device=device, requires_grad=False)
        t._node = node
        t._kind = kind
        return t

    def materialize(self) -> torch.Tensor:
        return self._node.value(self._kind).reshape(tuple(self.shape))

    @classmethod
    def __torch_dispatch__(cls, func, types, args=(), kwargs=None):
        return func(*tree_map(_real, args), **tree_map(_real, kwargs or {}))

    @classmethod
    def __torch_function__(cls, func, types, args=(), kwargs=None):
        kwargs = kwargs or {}
        name = getattr(func, "__name__", "")
        if name == "__get__" or name in _META:
            with torch._C.DisableTorchFunctionSubclass():
                return func(*args, **kwargs)
        if func in (torch.Tensor.reshape, torch.reshape, torch.Tensor.view) and isinstance(args[0], Deferred) and not kwargs:
            new = _resolve_shape(args[0], args[1:])
            if new is not None:
                return Deferred(new, args[0].device, args[0]._node, args[0]._kind)
        if func in (torch.Tensor.contiguous, torch.Tensor.float, torch.Tensor.detach) and isinstance(args[0], Deferred):
            return args[0]
        with torch._C.DisableTorchFunctionSubclass():
            return func(*tree_map(_real, args), **tree_map(_real, kwargs))


def _real(x):
    return x.materialize() if isinstance(x, Deferred) else x


def _resolve_shape(t: Deferred, shape_args):
    """New shape of a reshape/view that keeps the trailing (channel) dimension; None if it does not."""
    shp = shape_args[0] if len(shape_args) == 1 and isinstance(shape_args[0], (tuple, list, torch.Size)) else shape_args
    try:
        shp = [int(v) for v in shp]
    except (TypeError, ValueError):
        return None
    total = 1
    for v in t.shape:
        total *= int(v)
    if shp.count(-1) > 1:
        return None
    if -1 in shp:
        known = 1
        for v in shp:
            if v != -1:
                known *= v
        if known == 0 or total % known:
            return None
        shp[shp.index(-1)] = total // known
    prod = 1
    for v in shp:
        prod *= v
    if prod != total or not shp or shp[-1] != int(t.shape[-1]):
        return None
    return shp


def make_points(spec: SampleSpec):
    node = Node(spec)
    if not FUSE:
        return node.value("pts")
    return Deferred((spec.n, spec.S, 3), spec.rd.device, node, "pts")


def same_tensor(a: torch.Tensor, b: torch.Tensor) -> bool:
    return (a is b) or (a.data_ptr() == b.data_ptr() and a.shape == b.shape and a.stride() == b.stride()
                        and a.dtype == b.dtype)
