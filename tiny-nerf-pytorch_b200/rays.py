"""rays.get_rays -- drop-in for the reference's src/rays.py:3-33, computed by tnerf_get_rays."""
import torch

import _engine as E


def get_rays(H: int, W: int, focal: float, c2w: torch.Tensor, device=None):
    """Ray origins / unit directions of one HxW pinhole camera (camera looks along -z).

    Same contract as the reference: returns ``rays_o`` as a broadcast (stride-0) view of
    ``c2w[:3, 3]`` and ``rays_d`` as a dense (H*W, 3) fp32 tensor, ray k = row*W + col.
    """
    device = torch.device(device) if device is not None else c2w.device
    pose = E.f32c(c2w.to(device))
    E.need_cuda(pose)
    n = int(H) * int(W)
    rays_d = torch.empty((n, 3), dtype=torch.float32, device=device)
    E.check(E.lib().tnerf_get_rays(int(H), int(W), float(focal), E.ptr(pose), 0, n, None, E.ptr(rays_d),
                                   E.stream(device)), "tnerf_get_rays")
    rays_o = pose[:3, 3].expand_as(rays_d)
    return rays_o, rays_d
