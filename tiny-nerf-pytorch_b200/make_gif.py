"""make_gif.py -- novel-view GIF from the latest checkpoint (reference: src/make_gif.py:9-31).
The whole camera path is rendered by ONE pose-batched launch of the fused kernel (engine.render_frames); the reference's
own make_gif.py (render_one per pose) also runs unchanged against these modules."""
import os

import numpy as np
import torch

from _compat import imageio_v2
from camera import spiral_poses
from data import load_tiny_nerf_npz
from encoding import PositionalEncoding
from nerf import TinyNeRF
import engine


def main(ckpt_path="checkpoints/tinynerf_latest.pth", out="outputs/novel_views.gif", n_frames=60, radius=0.3):
    imageio = imageio_v2()
    device = torch.device("cuda")
    blob = load_tiny_nerf_npz("data/tiny_nerf_data.npz")
    poses = torch.from_numpy(blob["poses"]).to(device)
    _, H, W, _ = blob["images"].shape
    focal = float(blob["focal"])
    encoder = PositionalEncoding(num_freqs=10, include_input=True).to(device)
    ckpt = torch.load(ckpt_path, map_location=device)
    model = TinyNeRF(in_dim=encoder.out_dim, **ckpt.get("cfg", dict(hidden=128, depth=4, skip_at=2))).to(device)
    model.load_state_dict(ckpt["model"])
    path = spiral_poses(poses[0], n_frames=n_frames, radius=radius).to(device)
    imgs = engine.render_frames(model, encoder, H, W, focal, path, n_samples=64, near=2.0, far=6.0)
    frames = list((imgs.cpu().numpy() * 255).astype(np.uint8))
    print(f"[render] {n_frames}/{n_frames}")
    os.makedirs(os.path.dirname(out) or ".", exist_ok=True)
    imageio.mimsave(out, frames, fps=15, loop=0)
    print(f"[ok] wrote {out}")


if __name__ == "__main__":
    main()
