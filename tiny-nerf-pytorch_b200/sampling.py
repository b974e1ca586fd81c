"""sampling.stratified_samples -- drop-in for the reference's src/sampling.py:3-28."""
import torch

import _engine as E
import _lazy
import engine


def _as_per_ray(v, n, device):
    """near/far may be floats or tensors broadcastable to (N_rays, 1) (src/sampling.py:8)."""
    if torch.is_tensor(v):
        if v.numel() == 1:
            return float(v), None
        return 0.0, E.f32c(v.to(device).expand(n, 1).reshape(n))
    return float(v), None


def stratified_samples(near, far, n_samples, rays_o, rays_d, randomized=True, t_rand=None):
    """Depths ``z_vals (N,S)`` and points ``pts (N,S,3)`` along each ray.

    ``t_rand`` is an additive extension: an explicit uniform jitter tensor (N,S).  Without it the
    jitter is drawn from the device's global generator exactly where the reference calls
    ``torch.rand_like`` (src/sampling.py:24).  As in the reference, the non-randomised ``z_vals`` is an
    expanded view of one row.

    ``pts`` is returned as a deferred tensor (see _lazy.py): it behaves like the (N,S,3) tensor, but if
    it only flows through PositionalEncoding -> TinyNeRF -> volume_render (src/train.py:114-121,
    :51-56) the whole chain runs as one fused kernel and the points are never written to HBM.
    """
    dev = E.need_cuda(rays_o, rays_d)
    n, S = int(rays_o.shape[0]), int(n_samples)
    rd = E.f32c(rays_d)
    ro, o_stride = engine.origin_arg(rays_o)
    nr, nr_t = _as_per_ray(near, n, dev)
    fr, fr_t = _as_per_ray(far, n, dev)
    jitter = None
    if randomized:
        jitter = E.f32c(t_rand.to(dev)) if t_rand is not None else torch.rand((n, S), dtype=torch.float32, device=dev)
        if tuple(jitter.shape) != (n, S):
            raise ValueError(f"t_rand must have shape {(n, S)}")
    per_ray = nr_t is not None or fr_t is not None
    if per_ray:
        nr_t = nr_t if nr_t is not None else torch.full((n,), nr, dtype=torch.float32, device=dev)
        fr_t = fr_t if fr_t is not None else torch.full((n,), fr, dtype=torch.float32, device=dev)
    rows = n if (randomized or per_ray) else 1
    z = torch.empty((rows, S), dtype=torch.float32, device=dev)
    E.check(E.lib().tnerf_stratified(E.ptr(ro), o_stride, E.ptr(rd), rows, S, nr, fr, E.ptr(nr_t), E.ptr(fr_t),
                                     E.ptr(jitter), E.ptr(z), None, E.stream(dev)), "tnerf_stratified")
    z_vals = z if rows == n else z.expand(n, S)
    spec = _lazy.SampleSpec(ro=ro, o_stride=o_stride, rd=rd, n=n, S=S, near=nr, far=fr, near_t=nr_t, far_t=fr_t,
                            jitter=jitter, z_vals=z_vals)
    return z_vals, _lazy.make_points(spec)
