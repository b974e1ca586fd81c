"""main.py -- smoke render of pose 0 with an untrained network, the reference's src/main.py:14-62
(same prints, same output file) on the fused engine."""
import os
import time

import numpy as np
import torch

import engine
from _compat import imageio_v2
from data import load_tiny_nerf_npz
from encoding import PositionalEncoding
from nerf import TinyNeRF
from rays import get_rays


@torch.no_grad()
def test_render_once(model, encoder, H, W, focal, pose, device, n_samples=64, near=2.0, far=6.0, chunk=8192):
    """(H,W,3) image of one pose; one fused launch (`chunk` kept for signature compatibility)."""
    model.eval()
    # rays are generated inside the render kernel from the pose (a1 fused in)
    return engine.render_frames(model, encoder, H, W, focal, pose.to(device).reshape(1, 4, 4), n_samples=n_samples, near=near, far=far)[0]


def main():
    imageio = imageio_v2()
    torch.manual_seed(0)
    np.random.seed(0)
    if not torch.cuda.is_available():
        raise RuntimeError("this engine runs on CUDA (sm_100a) only")
    device = torch.device("cuda")
    print(f"[device] {device} torch={torch.__version__}")
    blob = load_tiny_nerf_npz("data/tiny_nerf_data.npz")
    images, poses, focal = torch.from_numpy(blob["images"]), torch.from_numpy(blob["poses"]), float(blob["focal"])
    N, H, W, _ = images.shape
    print(f"[data] N={N} H={H} W={W} focal={focal:.2f}")
    encoder = PositionalEncoding(num_freqs=10, include_input=True).to(device)
    model = TinyNeRF(in_dim=encoder.out_dim, hidden=128, depth=4, skip_at=2).to(device)
    os.makedirs("outputs", exist_ok=True)
    t0 = time.time()
    img = test_render_once(model, encoder, H, W, focal, poses[0], device)
    frame = (img.cpu().numpy() * 255).astype(np.uint8)          # the copy back synchronises, so dt is a real time
    dt = time.time() - t0
    imageio.imwrite("outputs/preview.png", frame)
    print(f"[render] wrote outputs/preview.png in {dt:.2f}s (untrained model; expect noisy image)")


if __name__ == "__main__":
    main()
