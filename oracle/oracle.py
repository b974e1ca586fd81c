"""CPU oracle for the TinyNeRF ray hot path.  TEST INFRASTRUCTURE ONLY.

This module is a restatement, in plain torch-on-CPU fp32 (or fp64) arithmetic, of the
algorithm the reference (avihaig/tiny-nerf-pytorch) runs for

    get_rays -> stratified_samples -> PositionalEncoding -> TinyNeRF -> volume_render
    (+ MSE loss, backward, Adam)

It exists only to CHECK the CUDA product path.  Only ``tests/``, ``__graft_entry__.smoke()``
and the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it.  Nothing under
``tiny-nerf-pytorch_b200/`` imports it; the product path has no CPU fallback.

Parity status: the reference ships no tests, golden vectors or fixtures for this path
(SURVEY.md section 4), so the oracle is pinned against OUTPUTS OF THE REFERENCE ITSELF: the script
``tests/golden/make_golden.py`` imports the unmodified reference modules from
``/root/reference/src`` in the build container and stores input/output vectors in
``tests/golden/*.npz``; ``tests/test_oracle_golden.py`` replays them through this file.
The one stage the reference delegates to a third-party generator -- the stratified jitter, ``torch.rand_like`` on the device,
src/sampling.py:24 -- is restated as Philox4x32-10 with the engine's documented keying and pinned against the published
Random123 known-answer vectors (bottom of this file).

Every function cites the reference lines it restates (paths relative to /root/reference).
The arithmetic is deliberately written in a different shape from the reference (functional,
explicit parameter dict, closed-form index maps) -- it follows the maths, not the text.
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Sequence, Tuple

import torch

Tensor = torch.Tensor
Params = Dict[str, Tensor]


# --------------------------------------------------------------------------------------
# a1  rays  (src/rays.py:14-33)
# --------------------------------------------------------------------------------------
def get_rays(H: int, W: int, focal: float, c2w: Tensor) -> Tuple[Tensor, Tensor]:
    """Pinhole rays for one pose.  Ray k = row*W + col (x fastest), camera looks down -z.

    src/rays.py:15-25 builds dirs = ((col - W/2)/focal, -(row - H/2)/focal, -1) from an int64
    pixel grid; :28-31 rotates by c2w[:3,:3] (as a matmul with R^T) and L2-normalises with
    eps 1e-12; :32 broadcasts the translation as the origin.
    """
    dt = c2w.dtype
    k = torch.arange(H * W, dtype=torch.int64)
    col = (k % W)
    row = (k // W)
    # int64 minus python float promotes to the default float dtype (fp32) exactly like the reference
    dx = (col - W * 0.5) / focal
    dy = -(row - H * 0.5) / focal
    dz = -torch.ones(H * W, dtype=torch.float32)
    cam = torch.stack([dx, dy, dz], dim=1).to(torch.float32).to(dt)
    world = cam @ c2w[:3, :3].transpose(0, 1)
    nrm = world.norm(dim=1, keepdim=True).clamp_min(1e-12)
    rays_d = world / nrm
    rays_o = c2w[:3, 3].unsqueeze(0).expand(H * W, 3)
    return rays_o, rays_d


# --------------------------------------------------------------------------------------
# a3  stratified samples  (src/sampling.py:14-28)
# --------------------------------------------------------------------------------------
def depth_bins(near, far, n_samples: int, dtype=torch.float32) -> Tensor:
    """z_i = near*(1-t_i) + far*t_i with t = linspace(0,1,S)   (src/sampling.py:16-17)."""
    t = torch.linspace(0.0, 1.0, steps=n_samples, dtype=dtype)
    return near * (1.0 - t) + far * t


def stratified(near, far, n_samples: int, rays_o: Tensor, rays_d: Tensor,
               u: Optional[Tensor] = None) -> Tuple[Tensor, Tensor]:
    """Depths and 3-D points.  ``u`` is the explicit uniform jitter tensor (N,S); None means the
    non-randomised path.  src/sampling.py:20-25: bin edges are the mid-points with the first and
    last bins clamped to z_0 / z_{S-1} (half width); z = lo + (hi-lo)*u.  :27: pts = o + d*z.
    """
    n = rays_o.shape[0]
    z = depth_bins(near, far, n_samples, rays_o.dtype)
    z = z.expand(n, n_samples)
    if u is not None:
        mid = (z[:, 1:] + z[:, :-1]) * 0.5
        hi = torch.cat([mid, z[:, -1:]], dim=1)
        lo = torch.cat([z[:, :1], mid], dim=1)
        z = lo + (hi - lo) * u
    pts = rays_o.unsqueeze(1) + rays_d.unsqueeze(1) * z.unsqueeze(2)
    return z, pts


# --------------------------------------------------------------------------------------
# a4  positional encoding  (src/encoding.py:14,26-33)
# --------------------------------------------------------------------------------------
def posenc(x: Tensor, num_freqs: int = 10, include_input: bool = True) -> Tensor:
    """[x, sin(2^0 x), cos(2^0 x), sin(2^1 x), ...]: column 3 + 6k + 3*{sin:0,cos:1} + axis."""
    if x.shape[-1] != 3:
        raise AssertionError("PositionalEncoding expects (..., 3)")
    parts = [x] if include_input else []
    for k in range(num_freqs):
        arg = x * float(2 ** k)
        parts += [arg.sin(), arg.cos()]
    return torch.cat(parts, dim=-1)


def posenc_dim(num_freqs: int, include_input: bool = True) -> int:
    return 6 * num_freqs + (3 if include_input else 0)


# --------------------------------------------------------------------------------------
# a5  MLP  (src/nerf.py:10-41)
# --------------------------------------------------------------------------------------
def mlp_param_shapes(in_dim: int, hidden: int = 128, depth: int = 4, skip_at: int = 2):
    """state_dict keys and shapes in construction order (src/nerf.py:19-27)."""
    shapes = []
    last = in_dim
    for i in range(depth):
        shapes.append((f"layers.{i}.weight", (hidden, last)))
        shapes.append((f"layers.{i}.bias", (hidden,)))
        last = hidden + in_dim if i == skip_at - 1 else hidden
    shapes += [("sigma.0.weight", (1, hidden)), ("sigma.0.bias", (1,)),
               ("rgb.0.weight", (3, hidden)), ("rgb.0.bias", (3,))]
    return shapes


def init_params(in_dim: int, hidden: int = 128, depth: int = 4, skip_at: int = 2,
                seed: int = 0, dtype=torch.float32) -> Params:
    """Deterministic U(-1/sqrt(fan_in), 1/sqrt(fan_in)) parameters (same family as nn.Linear's
    default; NOT the same RNG stream -- parity tests always share an explicit state_dict)."""
    g = torch.Generator().manual_seed(seed)
    out: Params = {}
    fan = None
    for name, shp in mlp_param_shapes(in_dim, hidden, depth, skip_at):
        if name.endswith("weight"):
            fan = shp[1]
        bound = 1.0 / math.sqrt(fan)
        out[name] = ((torch.rand(shp, generator=g, dtype=torch.float64) * 2 - 1) * bound).to(dtype)
    return out


def mlp_forward(p: Params, x: Tensor, depth: int = 4, skip_at: int = 2, return_pre: bool = False):
    """h = relu(W_i h + b_i); after layer skip_at-1 the input is appended BEHIND h
    (src/nerf.py:35-38); rgb = sigmoid(W_c h + b_c), sigma = relu(W_s h + b_s) (:26-27,39-41)."""
    h = x
    for i in range(depth):
        h = torch.addmm(p[f"layers.{i}.bias"], h, p[f"layers.{i}.weight"].t()).clamp_min(0)
        if i == skip_at - 1:
            h = torch.cat([h, x], dim=1)
    rgb = torch.sigmoid(torch.addmm(p["rgb.0.bias"], h, p["rgb.0.weight"].t()))
    pre = torch.addmm(p["sigma.0.bias"], h, p["sigma.0.weight"].t())
    sigma = pre.clamp_min(0)
    if return_pre:
        return rgb, sigma, pre
    return rgb, sigma


def last_sample_sigma_pre(p: Params, rays_o: Tensor, rays_d: Tensor, near, far, n_samples: int, u=None,
                          num_freqs: int = 10, include_input: bool = True, depth: int = 4, skip_at: int = 2) -> Tensor:
    """Pre-activation density of each ray's LAST sample.  Because delta_last = 1e10 (src/volume.py:20),
    a ray's acc/rgb/depth are discontinuous in this value at 0: parity tests use it to
    set aside rays whose last sample sits on the discontinuity."""
    z, pts = stratified(near, far, n_samples, rays_o, rays_d, u)
    feat = posenc(pts[:, -1, :], num_freqs, include_input)
    return mlp_forward(p, feat, depth, skip_at, return_pre=True)[2].reshape(-1)


# --------------------------------------------------------------------------------------
# a6  alpha compositing  (src/volume.py:18-44)
# --------------------------------------------------------------------------------------
def composite(rgb: Tensor, sigma: Tensor, z: Tensor, rays_d: Tensor, white_bkgd: bool = True):
    """delta_i = (z_{i+1}-z_i)*|d|, delta_last = 1e10*|d| (:18-23); alpha = 1-exp(-sigma*delta)
    (:27); T = exclusive cumprod of (1-alpha+1e-10) (:30-32); w = alpha*T (:34); sums (:36-38);
    white background adds 1-acc (:41-42).  Returns (rgb, depth, acc, weights) (:44).
    Deliberate deviation at n_samples = 1: the reference takes delta_last from `deltas[..., :1]` of an EMPTY tensor (:19-21) and ends
    up compositing no sample (weights (N, 0), background only); here the single sample is a last sample."""
    n, s = z.shape
    gap = torch.empty_like(z)
    gap[:, : s - 1] = z[:, 1:] - z[:, :-1]
    gap[:, s - 1] = 1e10
    gap = gap * rays_d.norm(dim=1, keepdim=True)
    alpha = 1.0 - torch.exp(-sigma.reshape(n, s) * gap)
    q = 1.0 - alpha + 1e-10
    run = torch.cumprod(q, dim=1)
    trans = torch.cat([torch.ones_like(run[:, :1]), run[:, :-1]], dim=1)
    w = alpha * trans
    acc = w.sum(dim=1, keepdim=True)
    col = (w.unsqueeze(2) * rgb).sum(dim=1)
    depth = (w * z).sum(dim=1, keepdim=True)
    if white_bkgd:
        col = col + (1.0 - acc)
    return col, depth, acc, w


def composite_backward(rgb, sigma, z, rays_d, gC, gD, gA, gW=None, white_bkgd=True):
    """Analytic backward of ``composite`` w.r.t. rgb and sigma (verified there
    against autograd): reverse scan R_{S-1}=0, dL/dalpha_i = T_i (g_i - R_i),
    R_{i-1} = g_i alpha_i + q_i R_i.  Used to validate the CUDA reverse scan independently of
    torch.autograd.  Inputs gC (N,3), gD (N,1), gA (N,1), gW (N,S) or None."""
    n, s = z.shape
    sig = sigma.reshape(n, s)
    gap = torch.empty_like(z)
    gap[:, : s - 1] = z[:, 1:] - z[:, :-1]
    gap[:, s - 1] = 1e10
    gap = gap * rays_d.norm(dim=1, keepdim=True)
    e = torch.exp(-sig * gap)
    alpha = 1.0 - e
    q = 1.0 - alpha + 1e-10
    run = torch.cumprod(q, dim=1)
    trans = torch.cat([torch.ones_like(run[:, :1]), run[:, :-1]], dim=1)
    w = alpha * trans
    g = (rgb * gC.unsqueeze(1)).sum(-1) + gD * z + gA
    if white_bkgd:
        g = g - gC.sum(-1, keepdim=True)
    if gW is not None:
        g = g + gW
    d_rgb = w.unsqueeze(2) * gC.unsqueeze(1)
    d_alpha = torch.zeros_like(z)
    R = torch.zeros(n, dtype=z.dtype)
    for i in range(s - 1, -1, -1):
        d_alpha[:, i] = trans[:, i] * (g[:, i] - R)
        R = g[:, i] * alpha[:, i] + q[:, i] * R
    d_sigma = d_alpha * gap * e
    return d_rgb, d_sigma.unsqueeze(-1)


# --------------------------------------------------------------------------------------
# a7  loss / metric  (src/train.py:122-123, src/utils.py:14-15)
# --------------------------------------------------------------------------------------
def mse(pred: Tensor, target: Tensor) -> Tensor:
    return ((pred - target) ** 2).mean()


def mse2psnr(m: Tensor) -> Tensor:
    return -10.0 * torch.log10(m.clamp_min(1e-10))


# --------------------------------------------------------------------------------------
# composed paths (src/train.py:46-58 render chunk body; :114-126 train step body)
# --------------------------------------------------------------------------------------
def render_rays(p: Params, rays_o: Tensor, rays_d: Tensor, near, far, n_samples: int,
                u: Optional[Tensor] = None, num_freqs: int = 10, include_input: bool = True,
                depth: int = 4, skip_at: int = 2, white_bkgd: bool = True):
    z, pts = stratified(near, far, n_samples, rays_o, rays_d, u)
    n = rays_o.shape[0]
    feat = posenc(pts.reshape(-1, 3), num_freqs, include_input)
    c, s = mlp_forward(p, feat, depth, skip_at)
    return composite(c.reshape(n, n_samples, 3), s.reshape(n, n_samples, 1), z, rays_d, white_bkgd)


def render_image(p: Params, H: int, W: int, focal: float, pose: Tensor, near=2.0, far=6.0,
                 n_samples: int = 64, chunk: int = 8192, **kw) -> Tensor:
    """Chunked full-frame render, clamp to [0,1]  (src/train.py:36-59)."""
    ro, rd = get_rays(H, W, focal, pose)
    out = []
    with torch.no_grad():
        for a in range(0, H * W, chunk):
            out.append(render_rays(p, ro[a:a + chunk], rd[a:a + chunk], near, far, n_samples, None, **kw)[0])
    return torch.cat(out, 0).reshape(H, W, 3).clamp(0.0, 1.0)


def loss_and_grads(p: Params, rays_o, rays_d, target, near, far, n_samples, u,
                   num_freqs=10, include_input=True, depth=4, skip_at=2, white_bkgd=True,
                   denom: Optional[int] = None):
    """MSE loss of the rendered colour against ``target`` and d(loss)/d(param) by autograd
    (src/train.py:114-126 on CPU where autocast/GradScaler are disabled).  ``denom`` overrides the
    3*N normaliser (ray-sharded data parallel uses the GLOBAL ray count)."""
    q = {k: v.detach().clone().requires_grad_(True) for k, v in p.items()}
    col, dep, acc, _ = render_rays(q, rays_o, rays_d, near, far, n_samples, u, num_freqs,
                                   include_input, depth, skip_at, white_bkgd)
    if denom is None:
        loss = mse(col, target)
    else:
        loss = ((col - target) ** 2).sum() / denom
    loss.backward()
    return loss.detach(), {k: v.grad for k, v in q.items()}, (col.detach(), dep.detach(), acc.detach())


def adam_step(p: Params, g: Params, m: Params, v: Params, step: int, lr=5e-4, b1=0.9, b2=0.999,
              eps=1e-8) -> None:
    """In-place torch.optim.Adam update, defaults of src/train.py:80 (no weight decay, no amsgrad).
    ``step`` is the 1-based step count AFTER this update."""
    c1 = 1.0 - b1 ** step
    c2 = 1.0 - b2 ** step
    for k in p:
        m[k].mul_(b1).add_(g[k], alpha=1 - b1)
        v[k].mul_(b2).addcmul_(g[k], g[k], value=1 - b2)
        denom = (v[k].sqrt() / math.sqrt(c2)).add_(eps)
        p[k].addcdiv_(m[k], denom, value=-lr / c1)


# --------------------------------------------------------------------------------------
# N3  camera path  (src/camera.py:4-12)
# --------------------------------------------------------------------------------------
def spiral_poses(c2w_ref: Tensor, n_frames: int = 60, radius: float = 0.3) -> Tensor:
    """c2w_ref @ Translate(r cos t, r sin t, 0) for t in linspace(0, 2pi, n)."""
    ts = torch.linspace(0, 2 * math.pi, n_frames)
    out = []
    for t in ts:
        T = torch.eye(4, dtype=c2w_ref.dtype)
        T[0, 3] = radius * torch.cos(t)
        T[1, 3] = radius * torch.sin(t)
        out.append(c2w_ref @ T)
    return torch.stack(out)


# --------------------------------------------------------------------------------------
# synthetic scene (the reference's dataset is absent: .MISSING_LARGE_BLOBS:1)
# --------------------------------------------------------------------------------------
def look_at_pose(theta: float, phi: float, radius: float = 4.0) -> Tensor:
    """Camera on a sphere looking at the origin; columns = right, up, back (camera looks down -z)."""
    eye = torch.tensor([radius * math.cos(phi) * math.cos(theta),
                        radius * math.cos(phi) * math.sin(theta),
                        radius * math.sin(phi)], dtype=torch.float64)
    back = eye / eye.norm()
    up0 = torch.tensor([0.0, 0.0, 1.0], dtype=torch.float64)
    right = torch.linalg.cross(up0, back)
    right = right / right.norm()
    up = torch.linalg.cross(back, right)
    m = torch.eye(4, dtype=torch.float64)
    m[:3, 0], m[:3, 1], m[:3, 2], m[:3, 3] = right, up, back, eye
    return m.to(torch.float32)


def analytic_field(pts: Tensor) -> Tuple[Tensor, Tensor]:
    """A smooth blob scene: density of three Gaussian lobes, colour varying with position."""
    centres = torch.tensor([[0.0, 0.0, 0.0], [0.7, 0.3, 0.2], [-0.5, -0.4, 0.4]], dtype=pts.dtype)
    widths = torch.tensor([0.55, 0.35, 0.3], dtype=pts.dtype)
    amps = torch.tensor([9.0, 14.0, 12.0], dtype=pts.dtype)
    d2 = ((pts.unsqueeze(-2) - centres) ** 2).sum(-1)
    sigma = (amps * torch.exp(-d2 / (2 * widths ** 2))).sum(-1, keepdim=True)
    col = torch.sigmoid(torch.stack([3 * pts[..., 0], 3 * pts[..., 1] + 1, 2 * pts[..., 2] - 1], -1))
    return col, sigma


def synthetic_scene(n_views: int = 8, H: int = 100, W: int = 100, focal: float = 138.88888549804688,
                    n_samples: int = 96, seed: int = 0):
    """images (N,H,W,3) f32, poses (N,4,4) f32, focal -- same keys/dtypes as tiny_nerf_data.npz
    (src/data.py:9-12, src/train.py:70-74), rendered from ``analytic_field`` with ``composite``."""
    g = torch.Generator().manual_seed(seed)
    poses, images = [], []
    for i in range(n_views):
        th = 2 * math.pi * i / n_views + 0.1 * float(torch.rand((), generator=g))
        ph = 0.3 + 0.5 * float(torch.rand((), generator=g))
        pose = look_at_pose(th, ph, 4.0)
        ro, rd = get_rays(H, W, focal, pose)
        z, pts = stratified(2.0, 6.0, n_samples, ro, rd, None)
        c, s = analytic_field(pts)
        img = composite(c, s, z, rd, True)[0].reshape(H, W, 3).clamp(0, 1)
        poses.append(pose)
        images.append(img)
    return {"images": torch.stack(images).numpy(), "poses": torch.stack(poses).numpy(),
            "focal": torch.tensor(focal, dtype=torch.float32).numpy()}


# --------------------------------------------------------------------------------------
# test helper: the two branches of the delta_last = 1e10 discontinuity
# --------------------------------------------------------------------------------------
def render_rays_last_flipped(p: Params, rays_o: Tensor, rays_d: Tensor, near, far, n_samples: int, u=None,
                             num_freqs: int = 10, include_input: bool = True, depth: int = 4, skip_at: int = 2,
                             white_bkgd: bool = True):
    """render_rays with the density of every ray's LAST sample moved to the other side of 0: off (sigma = 0) where the network
    says on, fully on (alpha_last = 1) where it says off.  src/volume.py:20-23 multiplies sigma_last by 1e10, so a ray's
    colour/depth/acc jump by T_last * (...) when the pre-activation crosses 0; a low-precision evaluation may legitimately land
    on the other branch for a pre-activation within rounding of 0.  Parity tests accept a set-aside ray only if it matches one
    of the two branches."""
    n = rays_o.shape[0]
    z, pts = stratified(near, far, n_samples, rays_o, rays_d, u)
    feat = posenc(pts.reshape(-1, 3), num_freqs, include_input)
    c, s = mlp_forward(p, feat, depth, skip_at)
    s = s.reshape(n, n_samples, 1).clone()
    s[:, -1, 0] = torch.where(s[:, -1, 0] > 0, torch.zeros_like(s[:, -1, 0]), torch.ones_like(s[:, -1, 0]))
    return composite(c.reshape(n, n_samples, 3), s, z, rays_d, white_bkgd)


# --------------------------------------------------------------------------------------
# stratified jitter drawn in-kernel  (src/sampling.py:24: torch.rand_like on the device)
# --------------------------------------------------------------------------------------
# The reference draws its jitter from torch's device generator, whose stream depends on launch geometry and cannot be reproduced
# outside torch.  The engine draws from the same FAMILY of generator (Philox4x32-10, Salmon et al., "Parallel random numbers: as
# easy as 1, 2, 3", SC'11 -- third-party algorithm, not part of /root/reference) with its own documented keying
# (include/tnerf.h, tnerf_ray_source): key = (seed.lo, seed.hi ^ step.hi), counter = (sample, ray.lo, ray.hi, step.lo), first output
# word, top 24 bits -> [0, 1).  Restated here in numpy integer arithmetic and pinned against the published known-answer vectors of
# the Random123 distribution (tests/test_oracle_golden.py::test_philox_known_answers); tnerf_jitter_fill is compared with it bit for bit.
def philox4x32_10(counter, key):
    """counter: 4 arrays of uint32 (broadcastable), key: 2 arrays of uint32 -> the 4 output words (uint32 arrays)."""
    import numpy as np
    c = [np.asarray(x, dtype=np.uint64) & np.uint64(0xFFFFFFFF) for x in counter]
    c = list(np.broadcast_arrays(*c))
    k0, k1 = (int(x) & 0xFFFFFFFF for x in key)
    m0, m1, mask = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57), np.uint64(0xFFFFFFFF)
    for _ in range(10):
        p0, p1 = m0 * c[0], m1 * c[2]                       # 32 x 32 -> 64 bit products
        c = [(p1 >> np.uint64(32)) ^ c[1] ^ np.uint64(k0), p1 & mask, (p0 >> np.uint64(32)) ^ c[3] ^ np.uint64(k1), p0 & mask]
        k0, k1 = (k0 + 0x9E3779B9) & 0xFFFFFFFF, (k1 + 0xBB67AE85) & 0xFFFFFFFF
    return [x.astype(np.uint32) for x in c]


def jitter_uniform(seed: int, step: int, n_rays: int, n_samples: int, first_ray: int = 0) -> Tensor:
    """(n_rays, n_samples) fp32 uniform [0, 1): the numbers the training kernel draws for (jitter_seed, jitter_step)."""
    import numpy as np
    ray = np.arange(first_ray, first_ray + n_rays, dtype=np.uint64)[:, None]
    smp = np.arange(n_samples, dtype=np.uint64)[None, :]
    x = philox4x32_10((smp, ray & np.uint64(0xFFFFFFFF), ray >> np.uint64(32), step & 0xFFFFFFFF),
                      (seed & 0xFFFFFFFF, ((seed >> 32) ^ (step >> 32)) & 0xFFFFFFFF))[0]
    return torch.from_numpy((x >> np.uint32(8)).astype(np.float32) * np.float32(2.0 ** -24))
