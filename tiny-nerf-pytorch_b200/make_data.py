"""Synthetic stand-in for data/tiny_nerf_data.npz (the reference's dataset is not shipped:
.MISSING_LARGE_BLOBS:1, scripts/get_data.sh needs the network).  Same keys / dtypes / conventions
(images (N,H,W,3) f32 in [0,1], poses (N,4,4) f32 camera-to-world looking down -z, focal scalar),
rendered on the GPU with this repo's own ray / sampling / compositing kernels from an analytic field."""
import argparse
import math
import os

import numpy as np
import torch

from rays import get_rays
from sampling import stratified_samples
from volume import volume_render


def look_at(theta, phi, radius=4.0):
    eye = np.array([radius * math.cos(phi) * math.cos(theta), radius * math.cos(phi) * math.sin(theta), radius * math.sin(phi)])
    back = eye / np.linalg.norm(eye)
    right = np.cross([0.0, 0.0, 1.0], back); right /= np.linalg.norm(right)
    up = np.cross(back, right)
    m = np.eye(4); m[:3, 0], m[:3, 1], m[:3, 2], m[:3, 3] = right, up, back, eye
    return m.astype(np.float32)


def field(pts):
    centres = torch.tensor([[0.0, 0.0, 0.0], [0.7, 0.3, 0.2], [-0.5, -0.4, 0.4]], device=pts.device)
    widths = torch.tensor([0.55, 0.35, 0.3], device=pts.device)
    amps = torch.tensor([9.0, 14.0, 12.0], device=pts.device)
    d2 = ((pts.unsqueeze(-2) - centres) ** 2).sum(-1)
    sigma = (amps * torch.exp(-d2 / (2 * widths ** 2))).sum(-1, keepdim=True)
    rgb = torch.sigmoid(torch.stack([3 * pts[..., 0], 3 * pts[..., 1] + 1, 2 * pts[..., 2] - 1], -1))
    return rgb, sigma


@torch.no_grad()
def make_scene(n_views=106, H=100, W=100, focal=138.88888549804688, n_samples=128, seed=0, device="cuda"):
    rng = np.random.default_rng(seed)
    poses, images = [], []
    for i in range(n_views):
        pose = look_at(2 * math.pi * i / n_views + 0.1 * rng.random(), 0.2 + 0.7 * rng.random())
        ro, rd = get_rays(H, W, focal, torch.from_numpy(pose).to(device))
        z, pts = stratified_samples(2.0, 6.0, n_samples, ro, rd, randomized=False)
        rgb, sigma = field(pts + 0)
        img = volume_render(rgb.contiguous(), sigma.contiguous(), z, rd)[0].reshape(H, W, 3).clamp(0, 1)
        poses.append(pose); images.append(img.cpu().numpy())
    return {"images": np.stack(images).astype(np.float32), "poses": np.stack(poses).astype(np.float32),
            "focal": np.array(focal, dtype=np.float64)}


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="data/tiny_nerf_data.npz")
    ap.add_argument("--views", type=int, default=106)
    ap.add_argument("--size", type=int, default=100)
    a = ap.parse_args()
    os.makedirs(os.path.dirname(a.out) or ".", exist_ok=True)
    np.savez(a.out, **make_scene(a.views, a.size, a.size, 138.88888549804688 * a.size / 100))
    print(f"[data] wrote {a.out}")
