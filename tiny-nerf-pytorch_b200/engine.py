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
    return _PREC[os.environ.get("TNERF_BWD_PRECISION", os.environ.get("TNERF_PRECISION", "f16")).lower()]


def ray_source(rays_o=None, o_stride=3, rays_d=None, c2w=None, H=0, W=0, focal=0.0, pixel_index=None, first_ray=0, jitter_seed=0, jitter_step=0):
    rs = E.RaySource()
    rs.jitter_seed, rs.jitter_step = int(jitter_seed), int(jitter_step)
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
        ctx.save_for_backward(*params)      # backward recomputes from the parameters: autograd's in-place-modification check guards them
        return comp, depth, acc

    @staticmethod
    def backward(ctx, gC, gD, gA):
        module, ro, o_stride, rd, n, S, near, far, jitter, white, prec = ctx.args
        ctx.saved_tensors                   # raises like the reference's autograd graph if a parameter was modified in place since the forward
        dev = rd.device
        h = E.handle_for(module, dev)
        ps = h.bind()
        if prec == E.PREC_F16_TC:
            h.ensure_packed()
        grads = torch.zeros(h.param_count, dtype=torch.float32, device=dev)
        rs = ray_source(ro, o_stride, rd)
        gC, gD, gA = (E.f32c(g) if g is not None else None for g in (gC, gD, gA))
        # grad_scale = 0: the tensor-core path picks its power-of-two loss scale on the device from the largest upstream gradient
        E.check(E.lib().tnerf_render_bwd(h.h, C.byref(rs), n, near, far, S, E.ptr(jitter), int(white), prec, E.ptr(gC), E.ptr(gD),
                                         E.ptr(gA), None, 0.0, None, E.ptr(grads), E.stream(dev)), "tnerf_render_bwd")
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
    if module.hidden == 256:          # wide model (BASELINE config 4): CTA-pair kernel, whole 32-sample chunks per warp
        return h.fused_ok and S % 32 == 0 and S // g <= 4
    return h.fused_ok and module.hidden == 128 and S // g <= 8


def train_supported(module, encoder, S: int, device) -> bool:
    """the tensor-core fused backward covers the reference MLP (depth 4, skip after layer 1, hidden 128) and
    ray tiles of whole rays (n_samples divides 128); everything else takes the fp32 path"""
    return (fused_supported(module, encoder, S, device) and module.hidden == 128 and module.depth == 4 and module.skip_at == 2
            and 1 <= int(S) <= 128 and 128 % int(S) == 0)


def pick_precisions(model, encoder, S, device, precision=None):
    prec = _PREC[precision.lower()] if precision else default_precision()
    bprec = _PREC[precision.lower()] if precision else default_bwd_precision()
    if prec == E.PREC_F16_TC and not fused_supported(model, encoder, S, device):
        prec = E.PREC_F32_SIMT
    if bprec == E.PREC_F16_TC and not train_supported(model, encoder, S, device):
        bprec = E.PREC_F32_SIMT
    return prec, bprec


def render_rays(model, encoder, rays_o, rays_d, near: float, far: float, n_samples: int, t_rand: Optional[torch.Tensor] = None,
                white_bkgd: bool = True, precision: Optional[str] = None):
    """Fused a3->a6: returns (comp_rgb (N,3), depth (N,1), acc (N,1)).  Differentiable w.r.t. the MLP
    parameters.  ``t_rand`` (N,S) enables stratified jitter; None = deterministic depths."""
    dev = E.need_cuda(rays_o, rays_d)
    h = E.handle_for(model, dev)
    h.set_encoding(encoder.num_freqs, encoder.include_input)
    prec, bprec = pick_precisions(model, encoder, n_samples, dev, precision)
    ro, o_stride = origin_arg(rays_o)
    rd = E.f32c(rays_d)
    jit = E.f32c(t_rand) if t_rand is not None else None
    return _FusedRender.apply(model, ro, o_stride, rd, int(rd.shape[0]), int(n_samples), float(near), float(far), jit,
                              bool(white_bkgd), prec, bprec, *model._params())


@torch.no_grad()
def render_frames(model, encoder, H: int, W: int, focal: float, poses: torch.Tensor, n_samples: int = 64, near: float = 2.0,
                  far: float = 6.0, white_bkgd: bool = True, precision: Optional[str] = None, return_aux: bool = False):
    """Pose-batched full-frame rendering: the frame loop of src/make_gif.py:22-27 (one ``render_one`` per pose) as ONE
    tnerf_render_frames call -- a single kernel launch for the whole camera path on the tensor-core path.
    poses (n,4,4) -> images (n,H,W,3) clamped to [0,1] like ``render_one``; with ``return_aux`` also depth and acc (n,H,W,1)."""
    dev = E.need_cuda(poses)
    h = E.handle_for(model, dev)
    h.set_encoding(encoder.num_freqs, encoder.include_input)
    prec, _ = pick_precisions(model, encoder, n_samples, dev, precision)
    if prec == E.PREC_F16_TC:
        h.ensure_packed()
    else:
        h.bind()
    poses_d = E.f32c(poses.reshape(-1, 4, 4))
    n, hw = int(poses_d.shape[0]), int(H) * int(W)
    comp = torch.empty((n * hw, 3), dtype=torch.float32, device=dev)
    depth = torch.empty((n * hw, 1), dtype=torch.float32, device=dev) if return_aux else None
    acc = torch.empty((n * hw, 1), dtype=torch.float32, device=dev) if return_aux else None
    E.check(E.lib().tnerf_render_frames(h.h, E.ptr(poses_d), n, int(H), int(W), float(focal), 0, hw, float(near), float(far), int(n_samples),
                                        int(white_bkgd), prec, E.ptr(comp), E.ptr(depth), E.ptr(acc), E.stream(dev)), "tnerf_render_frames")
    img = comp.reshape(n, H, W, 3).clamp(0, 1)
    if return_aux:
        return img, depth.reshape(n, H, W, 1), acc.reshape(n, H, W, 1)
    return img


def render_weights(model, ro, o_stride, rd, n, S, near, far, jitter, white, prec):
    """The (N,S) compositing weights of a fused render (the 4th value of volume_render), recomputed on demand."""
    dev = rd.device
    h = E.handle_for(model, dev)
    if prec == E.PREC_F16_TC and model.hidden != 128:
        prec = E.PREC_F32_SIMT        # the wide kernel does not write per-sample weights
    if prec == E.PREC_F16_TC:
        h.ensure_packed()
    comp = torch.empty((n, 3), dtype=torch.float32, device=dev)
    w = torch.empty((n, S), dtype=torch.float32, device=dev)
    rs = ray_source(ro, o_stride, rd)
    E.check(E.lib().tnerf_render_fwd(h.h, C.byref(rs), n, near, far, S, E.ptr(jitter), int(white), prec, E.ptr(comp), None, None,
                                     E.ptr(w), None, E.stream(dev)), "tnerf_render_fwd")
    return w


# SM-speed-aware tile dealing (include/tnerf.h, tnerf_set_tile_order): the SMs of one GPU differ by a few per cent in speed, stably per
# device.  The first Trainer on a device times the CTAs of the training kernel on a balanced batch (a handful of launches, once per
# process and device) and hands the resulting permutation to every handle it trains: when the tiles of a step do not divide evenly
# over the (CTA, stream) pairs, the slowest SMs run the shorter allotments.  TNERF_TILE_ORDER=0 keeps the identity.
_TILE_ORDER = os.environ.get("TNERF_TILE_ORDER", "1") != "0"
_TILE_ORDERS = {}        # device index -> (int32 device tensor: dealing index per CTA, summary dict)
_TILE_TABLES = []        # every table ever handed to a handle stays allocated (handles keep the bare pointer; a table is 592 bytes)


# One process on the tensor-core path: the gradient stays in the training kernel's sum vector (tnerf_train_fwd_bwd with grads = NULL)
# and the optimiser launch gathers it from there, transposing the weight blocks through shared memory -- the gradient-scatter launch
# disappears from the step (150.7 -> 148.4 us).  TNERF_GATHER=0: always scatter into the flat vector (developer A/B).
_GATHER = os.environ.get("TNERF_GATHER", "1") != "0"


class Trainer:
    """The training-step host path (src/train.py:106-128) on the fused kernels: one tnerf_train_fwd_bwd call (rays generated
    in-kernel from pose + pixel ids, MSE inside; training kernel + gradient scatter) and ONE optimiser launch (Adam + clearing
    of the gradient vector + in-place refresh of the fp16 operand image; with several ranks the same launch first all-reduces
    the gradient over NVLink peer memory).  No host synchronisation anywhere in ``step``.

    ``grad_scaler`` (default on) carries torch.amp.GradScaler's semantics (src/train.py:81,126-128) as device-resident state:
    a step whose scaled gradients overflow (or whose loss is not finite) is skipped -- parameters, moments and Adam's step
    count untouched -- and the loss scale is halved; after ``growth_interval`` clean steps it is doubled.

    Parameters stay ordinary ``nn.Parameter``s: they are re-pointed at views of one flat fp32 buffer so the optimiser is a
    single kernel; ``state_dict()`` / ``load_state_dict()`` speak torch.optim.Adam's format, so checkpoints interchange with
    the reference (src/train.py:85-92,142-156)."""

    def __init__(self, model, encoder, lr=5e-4, betas=(0.9, 0.999), eps=1e-8, near=2.0, far=6.0, n_samples=64,
                 white_bkgd=True, precision: Optional[str] = None, process_group=None, comm: Optional[str] = None,
                 grad_scaler: bool = True, init_scale: Optional[float] = None, growth_factor: float = 2.0, backoff_factor: float = 0.5,
                 growth_interval: int = 2000, jitter_seed: Optional[int] = None):
        self.model, self.encoder = model, encoder
        ps = model._params()
        dev = E.need_cuda(*ps)
        self.device = dev
        self.h = E.handle_for(model, dev)
        self.h.set_encoding(encoder.num_freqs, encoder.include_input)
        self.P = sum(p.numel() for p in ps)
        self.flat = torch.empty(self.P, dtype=torch.float32, device=dev)
        off = 0
        for p in ps:
            n = p.numel()
            self.flat[off:off + n].copy_(p.detach().reshape(-1))
            p.data = self.flat[off:off + n].view_as(p)
            off += n
        # [gradient | loss | overflow flag (even calls) | overflow flag (odd calls)]: one buffer = one all-reduce
        self.gbuf = torch.zeros(self.P + 3, dtype=torch.float32, device=dev)
        self.loss_view = self.gbuf[self.P:self.P + 1]
        self.loss_out = torch.zeros(1, dtype=torch.float32, device=dev)
        self.exp_avg = torch.zeros(self.P, dtype=torch.float32, device=dev)
        self.exp_avg_sq = torch.zeros(self.P, dtype=torch.float32, device=dev)
        self.lr, self.betas, self.eps = float(lr), (float(betas[0]), float(betas[1])), float(eps)
        self.near, self.far, self.S, self.white = float(near), float(far), int(n_samples), bool(white_bkgd)
        self.prec = pick_precisions(model, encoder, self.S, dev, precision)[1]
        self.pg = process_group
        self.world = 1
        if process_group is not None or (torch.distributed.is_available() and torch.distributed.is_initialized()):
            self.world = torch.distributed.get_world_size(process_group)
        # stratified jitter is drawn IN the training kernel (counter-based Philox keyed by (seed, step), include/tnerf.h) unless a
        # jitter tensor is passed to step_*(): the seed follows torch's global seed and differs per rank
        if jitter_seed is None:
            rk = torch.distributed.get_rank(process_group) if self.world > 1 else 0
            jitter_seed = (torch.initial_seed() * 0x9E3779B97F4A7C15 + rk * 0xD1B54A32D192ED03 + 0x632BE59BD9B4E019) & (2 ** 64 - 1)
        self.jitter_seed = int(jitter_seed) or 1
        self.steps = 0               # optimiser CALLS (skipped ones included); the applied count lives on the device with the scaler
        # GradScaler state (include/tnerf.h, tnerf_scaler): [scale, clean steps, applied steps, ..., 4 doubles of beta powers]
        self.scaler_state = torch.zeros(16, dtype=torch.float32, device=dev) if grad_scaler else None
        self._scale_init = init_scale
        self._scaler_cfg = (float(growth_factor), float(backoff_factor), int(growth_interval))
        if grad_scaler:
            self._reset_beta_powers(0)
            if init_scale is not None:
                self.scaler_state[0] = float(init_scale)
        self.h.bind()
        E.check(E.lib().tnerf_set_sum_buffer(self.h.h, None), "tnerf_set_sum_buffer")     # (a multi-rank trainer of the same model may have lent it one)
        if self.prec == E.PREC_F16_TC:
            self.h.ensure_packed(force=True)
        # gradient exchange of ray-sharded data parallel: "p2p" = one kernel that all-reduces over NVLink peer memory and
        # applies Adam (tnerf_allreduce_adam_step); "nccl" = torch.distributed.all_reduce + tnerf_optimizer_step
        self.comm = "none"
        if self.world > 1:
            want = comm or os.environ.get("TNERF_COMM", "p2p")
            self.comm = "nccl"
            if want == "p2p":
                try:
                    self._setup_p2p()
                    self.comm = "p2p"
                except Exception as e:  # noqa: BLE001  (no peer access / symmetric memory unavailable: NCCL does the exchange)
                    if comm == "p2p":
                        raise
                    print(f"[tnerf] peer-memory gradient exchange unavailable ({type(e).__name__}: {e}); using NCCL", flush=True)

        self.tile_order = None
        if _TILE_ORDER and self.prec == E.PREC_F16_TC:
            try:
                self.calibrate_tile_order()
            except Exception as e:  # noqa: BLE001  (an optimisation only: the identity dealing is always correct)
                print(f"[tnerf] tile-order calibration failed ({type(e).__name__}: {e}); keeping the identity dealing", flush=True)
                E.lib().tnerf_set_tile_order(self.h.h, None, 0)
                self.tile_order = None

    # ---- SM-speed-aware tile dealing ---------------------------------------------------------------
    def calibrate_tile_order(self, reps: int = 3, force: bool = False):
        """time the training kernel's CTAs on a balanced batch (every (CTA, stream) pair gets the same number of tiles) and give the
        handle the dealing order fastest SM first; cached per device.  Returns the summary (CTA loop times in ns)."""
        dev = self.device
        key = dev.index if dev.index is not None else torch.cuda.current_device()
        lib = E.lib()
        if force or key not in _TILE_ORDERS:
            sms = int(torch.cuda.get_device_properties(dev).multi_processor_count)
            tiles = 2 * sms * 8
            n = tiles * (64 // self.S) if self.S <= 64 else tiles // 2
            pose = torch.eye(4, device=dev)
            pose[2, 3] = 4.0
            pix = torch.arange(n, device=dev, dtype=torch.int64) % 10000
            tgt = torch.full((n, 3), 0.5, dtype=torch.float32, device=dev)
            scratch = torch.zeros(self.P + 3, dtype=torch.float32, device=dev)
            loss_slot = scratch[self.P:self.P + 1]
            dbg = torch.zeros(2048, dtype=torch.int64, device=dev)
            rs = ray_source(c2w=pose, H=100, W=100, focal=138.9, pixel_index=pix, jitter_seed=1)
            total = torch.zeros(sms, dtype=torch.float64)
            E.check(lib.tnerf_set_tile_order(self.h.h, None, 0), "tnerf_set_tile_order")
            E.check(lib.tnerf_set_debug_buffer(self.h.h, E.ptr(dbg)), "tnerf_set_debug_buffer")
            try:
                for r in range(reps + 1):                  # the first launch warms up
                    E.check(lib.tnerf_train_fwd_bwd(self.h.h, C.byref(rs), E.ptr(tgt), n, self.near, self.far, self.S, None, int(self.white),
                                                    self.prec, 3.0 * n, None, E.ptr(loss_slot), E.ptr(scratch), None, None, E.stream(dev)),
                            "tnerf_train_fwd_bwd (tile-order calibration)")
                    d = dbg.cpu()                          # synchronises
                    if r:
                        total += (d[1025:1025 + 4 * sms:4] - d[1024:1024 + 4 * sms:4]).double()
            finally:
                E.check(lib.tnerf_set_debug_buffer(self.h.h, None), "tnerf_set_debug_buffer")
            t = total / reps
            fastest_first = torch.argsort(t)
            perm = torch.empty(sms, dtype=torch.int32)
            perm[fastest_first] = torch.arange(sms, dtype=torch.int32)
            ts = t.sort().values
            summary = {"ctas": sms, "loop_ns_min": float(ts[0]), "loop_ns_median": float(ts[sms // 2]), "loop_ns_max": float(ts[-1]),
                       "slowest_ctas": [int(i) for i in fastest_first[-8:].flip(0)]}
            _TILE_ORDERS[key] = (perm.to(dev), summary)
            _TILE_TABLES.append(_TILE_ORDERS[key][0])
        table, summary = _TILE_ORDERS[key]
        self.tile_order = table                            # the handle keeps only the pointer: the tensor must outlive it
        E.check(lib.tnerf_set_tile_order(self.h.h, E.ptr(table), int(table.numel())), "tnerf_set_tile_order")
        return summary

    # ---- GradScaler state ------------------------------------------------------------------------
    def _reset_beta_powers(self, steps_done: int):
        pw = self.scaler_state[8:16].view(torch.float64)
        b1p, b2p = self.betas[0] ** steps_done, self.betas[1] ** steps_done
        pw.copy_(torch.tensor([b1p, b2p, b1p, b2p], dtype=torch.float64))
        self.scaler_state[2] = float(steps_done)

    @property
    def loss_scale(self) -> Optional[torch.Tensor]:
        """device tensor (1,) holding the current loss scale (None without the scaler)"""
        return None if self.scaler_state is None else self.scaler_state[0:1]

    def applied_steps(self) -> int:
        """optimiser steps actually applied (skipped steps do not count); synchronises"""
        return self.steps if self.scaler_state is None else int(self.scaler_state[2].item())

    def _scaler_struct(self, found, clear_next, call):
        if self.scaler_state is None:
            return None
        sc = E.Scaler()
        sc.state, sc.found_inf, sc.clear_next = E.ptr(self.scaler_state), E.ptr(found), E.ptr(clear_next)
        sc.growth_factor, sc.backoff_factor, sc.growth_interval = self._scaler_cfg
        sc.call = call
        return C.byref(sc)

    def _setup_p2p(self):
        """symmetric [grad | loss | overflow flag] buffers (double-buffered by step parity) + arrival flags, mapped into every rank"""
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm_mem
        group = self.pg if self.pg is not None else dist.group.WORLD
        self.rank = dist.get_rank(group)
        # sum-vector exchange (no gradient-scatter launch): the ranks' [sum | loss | flag] vectors in the training kernel's own order
        self._sum_elems = int(E.lib().tnerf_sum_elems(self.h.h)) if self.prec == E.PREC_F16_TC else -1
        stride = (max(self.P, self._sum_elems) + 2 + 63) // 64 * 64
        self._sym_stride = stride
        self.sym = symm_mem.empty(2 * stride + 64, dtype=torch.float32, device=self.device)
        self.sym.zero_()
        hdl = symm_mem.rendezvous(self.sym, group)
        torch.cuda.synchronize(self.device)
        dist.barrier(group)
        self._symh = hdl
        bases = [int(p) for p in hdl.buffer_ptrs]
        VP = C.c_void_p * self.world
        self._peer_grads = [VP(*[b + 4 * k * stride for b in bases]) for k in range(2)]
        self._peer_flags = VP(*[b + 4 * 2 * stride for b in bases])
        self._gviews = [self.sym[k * stride:k * stride + self.P + 2] for k in range(2)]
        self._sviews = [self.sym[k * stride:k * stride + self._sum_elems + 2] for k in range(2)] if self._sum_elems > 0 else None
        self.reduced = torch.zeros(self.P + 2, dtype=torch.float32, device=self.device)

    # ---- one optimisation step ------------------------------------------------------------------
    def _finish(self, gather=False):
        """optimiser step: ONE launch (Adam + clearing of the next step's gradient vector + in-place refresh of the fp16
        operand image); with several ranks the same launch first all-reduces the gradient over NVLink peer memory"""
        call = self.steps                # parity of THIS call: selects the overflow flag / beta-power slots
        self.steps += 1
        self.h.param_writes += 1         # (deferred outputs of earlier renders notice that the network has changed)
        st = E.stream(self.device)
        repack = 1 if self.prec == E.PREC_F16_TC else 0
        if not repack:
            # the optimiser kernel writes the flat parameters without touching their version counters: a later fp16 render
            # (previews, render_frames) must not reuse an operand image packed from older weights
            self.h.generation += 1
        P = self.P
        if self.comm == "p2p":
            k = self.steps & 1
            sc = self._scaler_struct(None, None, call)
            nxt = self._sviews[k ^ 1] if gather else self._gviews[k ^ 1]
            E.check(E.lib().tnerf_allreduce_adam_step(self.h.h, E.ptr(self.flat), E.ptr(self.exp_avg), E.ptr(self.exp_avg_sq), P,
                                                      self._peer_grads[k], self._peer_flags, self.world, self.rank, self.steps, self.steps,
                                                      self.lr, self.betas[0], self.betas[1], self.eps, E.ptr(self.reduced),
                                                      E.ptr(nxt), repack | (2 if gather else 0), sc, st), "tnerf_allreduce_adam_step")
            return self.reduced[P:P + 1]
        if self.world > 1:
            torch.distributed.all_reduce(self.gbuf, group=self.pg)
        # the gradient vector and its loss slot are cleared by the optimiser launch; the loss is handed out through loss_out
        sc = self._scaler_struct(self.gbuf[P + 1 + (call & 1):], self.gbuf[P + 1 + ((call + 1) & 1):], call)
        E.check(E.lib().tnerf_optimizer_step(self.h.h, E.ptr(self.flat), E.ptr(self.gbuf), E.ptr(self.exp_avg), E.ptr(self.exp_avg_sq), P,
                                             P + 1, self.steps, self.lr, self.betas[0], self.betas[1], self.eps, E.ptr(self.loss_out),
                                             repack | (2 if gather else 0), sc, st),
                "tnerf_optimizer_step")
        return self.loss_out

    def _launch(self, rs, target, n, jitter, global_rays):
        st = E.stream(self.device)
        P = self.P
        gbuf = self._gviews[(self.steps + 1) & 1] if self.comm == "p2p" else self.gbuf      # [gradient | loss | flag(s)] of this step, zero on entry
        denom = 3.0 * float(global_rays if global_rays else n * self.world)
        scale = found = None
        if self.scaler_state is not None:
            if self._scale_init is None:         # default initial scale: largest loss gradient 2/denom -> [64, 128) (fp16-friendly)
                self._scale_init = 2.0 ** (math.ceil(math.log2(denom)) + 6)
                self.scaler_state[0] = self._scale_init
            scale = self.scaler_state
            found = gbuf[P + 1:] if self.comm == "p2p" else gbuf[P + 1 + (self.steps & 1):]
        # one process, tensor-core path, one-vector flush: the gradient stays in the training kernel's sum vector and the optimiser
        # launch gathers it from there (no scatter launch); otherwise it is scattered into gbuf (the exchange vector of a multi-rank step)
        gather = self.prec == E.PREC_F16_TC and _GATHER and self.h.get_option("bulk_reduce") == 1 and \
            (self.world == 1 or (self.comm == "p2p" and self._sviews is not None))
        loss_slot = gbuf[P:]
        if gather and self.comm == "p2p":
            # several ranks: the kernel's sum vector IS this rank's exchange vector (peer-mapped), loss and overflow flag behind it
            sbuf = self._sviews[(self.steps + 1) & 1]
            E.check(E.lib().tnerf_set_sum_buffer(self.h.h, E.ptr(sbuf)), "tnerf_set_sum_buffer")
            loss_slot = sbuf[self._sum_elems:]
            if found is not None:
                found = sbuf[self._sum_elems + 1:]
        E.check(E.lib().tnerf_train_fwd_bwd(self.h.h, C.byref(rs), E.ptr(target), n, self.near, self.far, self.S, E.ptr(jitter),
                                            int(self.white), self.prec, denom, None, E.ptr(loss_slot), None if gather else E.ptr(gbuf),
                                            E.ptr(scale), E.ptr(found), st),
                "tnerf_train_fwd_bwd")
        return self._finish(gather)

    def step_pixels(self, c2w, H, W, focal, pixel_index, target, jitter=None, global_rays=None):
        """rays are generated in-kernel from the pose and the pixel ids (a1+a2 fused in); returns the loss (device, shape (1,)).
        ``jitter`` (n, n_samples) uniform [0,1) is the explicit stratified-jitter tensor (parity runs); None = drawn in-kernel."""
        n = int(pixel_index.shape[0])
        # the C ABI takes raw pointers: fp32 pose / targets / jitter, int64 pixel ids, all dense (no copies when they already are)
        E.need_cuda(c2w, pixel_index, target, jitter)
        c2w, target, jitter = E.f32c(c2w), E.f32c(target), E.f32c(jitter)
        if pixel_index.dtype != torch.int64 or not pixel_index.is_contiguous():
            pixel_index = pixel_index.long().contiguous()
        if tuple(target.shape) != (n, 3) or (jitter is not None and tuple(jitter.shape) != (n, self.S)) or c2w.numel() < 12:
            raise ValueError(f"step_pixels: expected target ({n}, 3), jitter ({n}, {self.S}), c2w (4, 4)")
        rs = ray_source(c2w=c2w, H=H, W=W, focal=focal, pixel_index=pixel_index, jitter_seed=self.jitter_seed, jitter_step=self.steps)
        return self._launch(rs, target, n, jitter, global_rays)

    def step_rays(self, rays_o, rays_d, target, jitter=None, global_rays=None):
        n = int(rays_d.shape[0])
        ro, o_stride = origin_arg(rays_o)
        rd = E.f32c(rays_d)
        tg = E.f32c(target)
        jitter = E.f32c(jitter)
        rs = ray_source(ro, o_stride, rd, jitter_seed=self.jitter_seed, jitter_step=self.steps)
        return self._launch(rs, tg, n, jitter, global_rays)

    def jitter_tensor(self, n: int, step: Optional[int] = None) -> torch.Tensor:
        """the (n, n_samples) jitter the kernel draws for optimiser call ``step`` (default: the next one) -- for parity runs"""
        out = torch.empty((n, self.S), dtype=torch.float32, device=self.device)
        E.check(E.lib().tnerf_jitter_fill(self.jitter_seed, self.steps if step is None else int(step), n, self.S, E.ptr(out), E.stream(self.device)),
                "tnerf_jitter_fill")
        return out

    # ---- torch.optim.Adam-compatible state ---------------------------------------------------------
    def state_dict(self):
        state, off = {}, 0
        applied = self.applied_steps()
        for i, p in enumerate(self.model._params()):
            n = p.numel()
            state[i] = {"step": torch.tensor(float(applied)), "exp_avg": self.exp_avg[off:off + n].view_as(p).clone(),
                        "exp_avg_sq": self.exp_avg_sq[off:off + n].view_as(p).clone()}
            off += n
        group = {"lr": self.lr, "betas": self.betas, "eps": self.eps, "weight_decay": 0, "amsgrad": False, "maximize": False,
                 "foreach": None, "capturable": False, "differentiable": False, "fused": None, "decoupled_weight_decay": False,
                 "params": list(range(len(state)))}
        return {"state": state if applied else {}, "param_groups": [group]}

    def load_state_dict(self, sd):
        off = 0
        for i, p in enumerate(self.model._params()):
            n = p.numel()
            st = sd["state"].get(i)
            if st is not None:
                self.exp_avg[off:off + n].copy_(st["exp_avg"].reshape(-1))
                self.exp_avg_sq[off:off + n].copy_(st["exp_avg_sq"].reshape(-1))
                self.steps = int(st["step"])
            off += n
        if sd.get("param_groups"):
            g = sd["param_groups"][0]
            self.lr, self.betas, self.eps = float(g["lr"]), tuple(g["betas"]), float(g["eps"])
        if self.scaler_state is not None:
            self._reset_beta_powers(self.steps)

    def refresh(self):
        """call after parameters were changed from outside (load_state_dict on the model)"""
        self.h.bind()
        if self.prec == E.PREC_F16_TC:
            self.h.ensure_packed(force=True)
