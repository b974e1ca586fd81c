"""ctypes binding of libtnerf.so (C ABI in include/tnerf.h) plus the small amount of host logic the
flat modules share: pointer/stream plumbing, the per-model handle, flat parameter storage.

There is deliberately no CPU path: a non-CUDA tensor reaching these helpers raises, and a missing
shared library raises at import of any op that needs it.
"""
from __future__ import annotations

import contextlib
import ctypes as C
import os
import weakref
from typing import List, Optional, Sequence

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("TNERF_LIB") or os.path.join(_HERE, "libtnerf.so")      # TNERF_LIB: developer override (A/B builds)

PREC_F16_TC = 0
PREC_F32_SIMT = 1

_p = C.c_void_p
_ll = C.c_longlong
_i = C.c_int
_f = C.c_float


class RaySource(C.Structure):
    """mirror of tnerf_ray_source (include/tnerf.h)"""
    _fields_ = [("rays_o", _p), ("o_stride", _ll), ("rays_d", _p), ("c2w", _p), ("H", _i), ("W", _i),
                ("focal", _f), ("pixel_index", _p), ("first_ray", _ll), ("jitter_seed", C.c_ulonglong), ("jitter_step", C.c_ulonglong)]


class Scaler(C.Structure):
    """mirror of tnerf_scaler (include/tnerf.h): device-resident GradScaler state consumed by the optimiser entry points"""
    _fields_ = [("state", _p), ("found_inf", _p), ("clear_next", _p), ("growth_factor", _f), ("backoff_factor", _f),
                ("growth_interval", _i), ("call", C.c_uint)]


ABI_VERSION = 2

_SIGS = {
    "tnerf_abi_version": (_i, []),
    "tnerf_last_error": (C.c_char_p, []),
    "tnerf_launch_count": (_ll, []),
    "tnerf_get_rays": (_i, [_i, _i, _f, _p, _ll, _ll, _p, _p, _p]),
    "tnerf_gather3": (_i, [_p, _ll, _ll, _p, _p, _p, _p, _p, _p, _p]),
    "tnerf_stratified": (_i, [_p, _ll, _p, _ll, _i, _f, _f, _p, _p, _p, _p, _p, _p]),
    "tnerf_posenc": (_i, [_p, _ll, _i, _i, _p, _p]),
    "tnerf_posenc_bwd": (_i, [_p, _p, _ll, _i, _i, _p, _p]),
    "tnerf_create": (_i, [C.POINTER(_p), _i, _i, _i, _i, _i]),
    "tnerf_destroy": (None, [_p]),
    "tnerf_bind_params": (_i, [_p, C.POINTER(_p), _i]),
    "tnerf_param_count": (_ll, [_p]),
    "tnerf_set_encoding": (_i, [_p, _i, _i]),
    "tnerf_set_option": (_i, [_p, C.c_char_p, _i]),
    "tnerf_get_option": (_i, [_p, C.c_char_p]),
    "tnerf_sum_elems": (_ll, [_p]),
    "tnerf_set_sum_buffer": (_i, [_p, _p]),
    "tnerf_clear_sum": (_i, [_p, _p]),
    "tnerf_set_debug_buffer": (_i, [_p, _p]),
    "tnerf_set_tile_order": (_i, [_p, _p, _i]),
    "tnerf_fused_supported": (_i, [_p]),
    "tnerf_pack_weights": (_i, [_p, _p]),
    "tnerf_mlp_fwd": (_i, [_p, _p, _ll, _p, _p, _p, _p]),
    "tnerf_mlp_bwd_scratch_floats": (_ll, [_p, _ll]),
    "tnerf_mlp_bwd": (_i, [_p, _p, _ll, _p, _p, _p, _p, _p, _p, _p, _p, _p]),
    "tnerf_composite_fwd": (_i, [_p, _p, _p, _ll, _p, _ll, _i, _i, _p, _p, _p, _p, _p]),
    "tnerf_composite_bwd": (_i, [_p, _p, _p, _ll, _p, _ll, _i, _i, _p, _p, _p, _p, _p, _p, _p]),
    "tnerf_render_fwd": (_i, [_p, C.POINTER(RaySource), _ll, _f, _f, _i, _p, _i, _i, _p, _p, _p, _p, _p, _p]),
    "tnerf_render_frames": (_i, [_p, _p, _i, _i, _i, _f, _ll, _ll, _f, _f, _i, _i, _i, _p, _p, _p, _p]),
    "tnerf_render_bwd": (_i, [_p, C.POINTER(RaySource), _ll, _f, _f, _i, _p, _i, _i, _p, _p, _p, _p, _f, _p, _p, _p]),
    "tnerf_train_fwd_bwd": (_i, [_p, C.POINTER(RaySource), _p, _ll, _f, _f, _i, _p, _i, _i, _f, _p, _p, _p, _p, _p, _p]),
    "tnerf_mse_psnr": (_i, [_p, _p, _ll, _p, _p]),
    "tnerf_adam_step": (_i, [_p, _p, _p, _p, _ll, _i, _f, _f, _f, _f, _f, _p, _p]),
    "tnerf_jitter_fill": (_i, [C.c_ulonglong, C.c_ulonglong, _ll, _i, _p, _p]),
    "tnerf_check_finite": (_i, [_p, _ll, _p, _p]),
    "tnerf_packed_image_copy": (_ll, [_p, _p, _ll, _p]),
    "tnerf_optimizer_step": (_i, [_p, _p, _p, _p, _p, _ll, _ll, _i, _f, _f, _f, _f, _p, _i, C.POINTER(Scaler), _p]),
    "tnerf_allreduce_adam_step": (_i, [_p, _p, _p, _p, _ll, _p, _p, _i, _i, C.c_uint, _i, _f, _f, _f, _f, _p, _p, _i, C.POINTER(Scaler), _p]),
    "tnerf_umma_rate": (_i, [_i, _i, _i, _p, _p]),
    "tnerf_umma_selftest": (_i, [_p, _p, _i, _i, _i, _p, _p]),
}

_lib = None


def lib():
    """The loaded C-ABI library; raises loudly when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(or `make -C tiny-nerf-pytorch_b200/csrc`). There is no fallback path.")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGS.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        if L.tnerf_abi_version() != ABI_VERSION:
            raise RuntimeError("libtnerf.so ABI version mismatch")
        _lib = L
    return _lib


def exported_symbols() -> List[str]:
    return sorted(_SIGS)


def check(rc: int, what: str = "tnerf") -> None:
    if rc != 0:
        msg = lib().tnerf_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"{what} failed (code {rc}): {msg}")


def need_cuda(*tensors) -> torch.device:
    dev = None
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise RuntimeError("tiny-nerf-pytorch_b200 runs on CUDA (sm_100a) only: got a "
                               f"{t.device} tensor; the CPU path is the reference implementation")
        dev = dev or t.device
    return dev


def f32c(t: Optional[torch.Tensor]) -> Optional[torch.Tensor]:
    """fp32 + contiguous (no copy when already so)"""
    if t is None:
        return None
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


def ptr(t: Optional[torch.Tensor]):
    return None if t is None else _p(t.data_ptr())


def stream(dev=None):
    return _p(torch.cuda.current_stream(dev).cuda_stream)


def launch_count() -> int:
    return int(lib().tnerf_launch_count())


# ------------------------------------------------------------------------------------------------
class ModelHandle:
    """tnerf_handle bound to one TinyNeRF module on one device.  Re-binds pointers when parameters
    move (``.to``) and re-packs the fp16 operand image when any parameter's version counter changes
    (the optimiser updates parameters in place)."""

    def __init__(self, module, device: torch.device):
        self.module = weakref.ref(module)
        self.device = device
        h = _p()
        check(lib().tnerf_create(C.byref(h), device.index if device.index is not None else torch.cuda.current_device(), module.in_dim,
                                 module.hidden, module.depth, module.skip_at), "tnerf_create")
        self.h = h
        self._bound = None
        self._packed_versions = None
        self.generation = 0          # bumped by writers that change the parameters without touching their version counters
        self.param_writes = 0        # every raw-kernel parameter update (engine.Trainer), whether or not the operand image was refreshed with it
        self.param_count = int(lib().tnerf_param_count(h))
        self.fused_ok = bool(lib().tnerf_fused_supported(h))

    def __del__(self):
        try:
            if self.h:
                lib().tnerf_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def params(self) -> List[torch.Tensor]:
        m = self.module()
        ps = []
        for lin in m.layers:
            ps += [lin.weight, lin.bias]
        ps += [m.sigma[0].weight, m.sigma[0].bias, m.rgb[0].weight, m.rgb[0].bias]
        return ps

    def bind(self) -> List[torch.Tensor]:
        ps = self.params()
        for p in ps:
            if p.dtype != torch.float32 or not p.is_contiguous():
                raise RuntimeError("TinyNeRF parameters must be contiguous fp32")
            need_cuda(p)
        key = tuple(p.data_ptr() for p in ps)
        if key != self._bound:
            arr = (_p * len(ps))(*[_p(k) for k in key])
            check(lib().tnerf_bind_params(self.h, arr, len(ps)), "tnerf_bind_params")
            self._bound = key
            self._packed_versions = None
        return ps

    def set_encoding(self, num_freqs: int, include_input: bool) -> None:
        check(lib().tnerf_set_encoding(self.h, int(num_freqs), int(bool(include_input))), "tnerf_set_encoding")
        self.fused_ok = bool(lib().tnerf_fused_supported(self.h))

    def set_option(self, name: str, value: int) -> None:
        """schedule options of the fused training kernel (include/tnerf.h, tnerf_set_option); -1 = built-in choice"""
        check(lib().tnerf_set_option(self.h, name.encode(), int(value)), "tnerf_set_option")

    def get_option(self, name: str) -> int:
        """the value in effect (defaults resolved); -1 = unknown option"""
        return int(lib().tnerf_get_option(self.h, name.encode()))

    def ensure_packed(self, force: bool = False) -> None:
        ps = self.bind()
        ver = tuple(p._version for p in ps) + (self.generation,)
        if force or ver != self._packed_versions:
            check(lib().tnerf_pack_weights(self.h, stream(self.device)), "tnerf_pack_weights")
            self._packed_versions = ver


# handles live OUTSIDE the module (keyed weakly by it): a ctypes pointer inside module.__dict__ would make
# copy.deepcopy(model) / torch.save(model) fail, which the reference nn.Module supports
_HANDLES: "weakref.WeakKeyDictionary" = weakref.WeakKeyDictionary()


def handle_for(module, device: torch.device) -> ModelHandle:
    cache = _HANDLES.get(module)
    if cache is None:
        cache = _HANDLES[module] = {}
    key = (device.type, device.index if device.index is not None else torch.cuda.current_device())
    h = cache.get(key)
    if h is None:
        h = ModelHandle(module, torch.device("cuda", key[1]))
        cache[key] = h
    return h


def on_device(dev):
    """context that makes ``dev`` the current CUDA device for a stand-alone launch (no-op when it already is): the kernels of the
    pointer-only entry points run on the current device, the reference's torch ops on the tensor's"""
    if dev is None or dev.index is None or dev.index == torch.cuda.current_device():
        return contextlib.nullcontext()
    return torch.cuda.device(dev)


def flat_grad_views(module, flat: torch.Tensor) -> List[torch.Tensor]:
    """Views of a flat (param_count,) gradient vector shaped like the parameters, state_dict order."""
    out, off = [], 0
    for p in handle_for(module, flat.device).params():
        n = p.numel()
        out.append(flat[off:off + n].view_as(p))
        off += n
    return out
