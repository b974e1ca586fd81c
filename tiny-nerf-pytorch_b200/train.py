"""train.py -- training driver with the reference's CLI (src/train.py:20-34,163), checkpoints and
previews, running on the fused B200 engine.

Behaviour kept from the reference: same `Config` fields/defaults (tyro CLI), seeds, img_i = step % N,
n_rand pixel ids drawn with torch.randint, Adam(lr), previews every `preview_every`, checkpoints
{"model","opt","step","in_dim","cfg"} every `ckpt_every` and at the end, resume by default, final.png of
the last pose.  What differs is HOW a step runs: rays are generated inside the fused kernel from the
pose and the pixel ids (no (N,HW,3) ray tables), forward+loss+backward are one launch, Adam is one
launch, and nothing synchronises with the host except the `log_every` read-out.

The reference's own src/train.py also runs unchanged against this directory's modules (its op-by-op
sequence is fused through deferred tensors, see _lazy.py); `--engine ops` runs that same sequence here.
"""
import os
import time
from dataclasses import dataclass
from typing import Optional

import numpy as np
import torch
from torch import nn

import engine
from _compat import imageio_v2
from data import load_tiny_nerf_npz
from encoding import PositionalEncoding
from nerf import TinyNeRF
from rays import get_rays
from sampling import stratified_samples
from utils import mse2psnr
from volume import volume_render

MODEL_CFG = dict(hidden=128, depth=4, skip_at=2)


@dataclass
class Config:
    iters: int = 20000          # optimisation steps
    n_rand: int = 2048          # rays (random pixels of one view) per step
    n_samples: int = 64         # samples per ray
    lr: float = 5e-4
    near: float = 2.0
    far: float = 6.0
    log_every: int = 50
    preview_every: int = 500
    ckpt_every: int = 1000
    ckpt_path: str = "checkpoints/tinynerf_latest.pth"
    out_dir: str = "outputs"
    resume: bool = True
    preview_pose: Optional[int] = None   # None -> (img_i + 1) % N
    engine: str = "fused"       # "fused": one-launch training step; "ops": the reference's op sequence + torch.optim.Adam
    data: str = "data/tiny_nerf_data.npz"


@torch.no_grad()
def render_one(model: nn.Module, encoder: nn.Module, H: int, W: int, focal: float, pose: torch.Tensor, device: torch.device,
               n_samples: int = 64, near: float = 2.0, far: float = 6.0, chunk: int = 8192) -> torch.Tensor:
    """Full (H,W,3) frame for one pose, clamped to [0,1].  `chunk` is accepted for signature compatibility:
    the fused kernel keeps no per-sample tensors in HBM, so the frame is rendered in one launch."""
    model.eval()
    # rays are generated inside the render kernel from the pose (a1 fused in): no get_rays launch, no (H*W,3) ray tensors
    return engine.render_frames(model, encoder, H, W, focal, pose.to(device).reshape(1, 4, 4), n_samples=n_samples, near=near, far=far)[0]


def _to_png(img: torch.Tensor) -> np.ndarray:
    return (img.cpu().numpy() * 255).astype(np.uint8)          # truncation, as the reference


def _checkpoint(path, model, opt_state, step, in_dim):
    torch.save({"model": model.state_dict(), "opt": opt_state, "step": step, "in_dim": in_dim, "cfg": dict(MODEL_CFG)}, path)


def main(cfg: Config):
    imageio = imageio_v2()
    torch.manual_seed(0)
    np.random.seed(0)
    if not torch.cuda.is_available():
        raise RuntimeError("this engine runs on CUDA (sm_100a) only; the CPU path is the reference implementation")
    device = torch.device("cuda")
    os.makedirs(cfg.out_dir, exist_ok=True)
    os.makedirs(os.path.dirname(cfg.ckpt_path) or ".", exist_ok=True)
    print(f"[device] {device} torch={torch.__version__}")

    blob = load_tiny_nerf_npz(cfg.data)
    images = torch.from_numpy(blob["images"]).to(device)
    poses = torch.from_numpy(blob["poses"]).to(device)
    focal = float(blob["focal"])
    N, H, W, _ = images.shape
    pixels = images.view(N, H * W, 3)
    print(f"[data] N={N} H={H} W={W} focal={focal:.2f}")

    encoder = PositionalEncoding(num_freqs=10, include_input=True).to(device)
    model = TinyNeRF(in_dim=encoder.out_dim, **MODEL_CFG).to(device)
    fused = cfg.engine == "fused"
    if fused:
        trainer = engine.Trainer(model, encoder, lr=cfg.lr, near=cfg.near, far=cfg.far, n_samples=cfg.n_samples)
        optimizer, scaler = trainer, None
    else:
        optimizer = torch.optim.Adam(model.parameters(), lr=cfg.lr)
        scaler = torch.amp.GradScaler("cuda")
        rays_tab = [get_rays(H, W, focal, poses[i], device=device) for i in range(N)]

    start = 0
    if cfg.resume and os.path.exists(cfg.ckpt_path):
        ck = torch.load(cfg.ckpt_path, map_location=device)
        model.load_state_dict(ck["model"])
        if "opt" in ck:
            optimizer.load_state_dict(ck["opt"])
        start = int(ck.get("step", 0))
        if fused:
            trainer.refresh()
        print(f"[resume] loaded {cfg.ckpt_path} from step {start}")

    try:
        from tqdm import tqdm
        bar = tqdm(range(start, cfg.iters), desc="train")
    except ImportError:
        bar = range(start, cfg.iters)
    t_begin = time.time()
    for step in bar:
        model.train()
        view = step % N
        pick = torch.randint(0, H * W, (cfg.n_rand,), device=device)
        if fused:
            target = pixels[view].index_select(0, pick)
            loss = trainer.step_pixels(poses[view], H, W, focal, pick, target)
        else:
            ro, rd = rays_tab[view][0][pick], rays_tab[view][1][pick]
            target = pixels[view, pick]
            z_vals, pts = stratified_samples(cfg.near, cfg.far, cfg.n_samples, ro, rd, randomized=True)
            with torch.amp.autocast("cuda"):
                rgb, sigma = model(encoder(pts.reshape(-1, 3)))
                comp, _, _, _ = volume_render(rgb.reshape(cfg.n_rand, cfg.n_samples, 3), sigma.reshape(cfg.n_rand, cfg.n_samples, 1), z_vals, rd)
                loss = torch.mean((comp - target) ** 2)
            optimizer.zero_grad(set_to_none=True)
            scaler.scale(loss).backward()
            scaler.step(optimizer)
            scaler.update()

        done = step + 1
        if done % cfg.log_every == 0 and hasattr(bar, "set_postfix"):
            lv = float(loss.reshape(-1)[0].item())
            bar.set_postfix(loss=lv, psnr=float(mse2psnr(torch.tensor(lv))))
        if done % cfg.preview_every == 0:
            idx = ((view + 1) if cfg.preview_pose is None else cfg.preview_pose) % N
            frame = render_one(model, encoder, H, W, focal, poses[idx], device, n_samples=cfg.n_samples, near=cfg.near, far=cfg.far)
            imageio.imwrite(f"{cfg.out_dir}/preview_{done:06d}.png", _to_png(frame))
        if done % cfg.ckpt_every == 0:
            _checkpoint(cfg.ckpt_path, model, optimizer.state_dict(), done, encoder.out_dim)

    torch.cuda.synchronize()
    minutes = (time.time() - t_begin) / 60
    _checkpoint(cfg.ckpt_path, model, optimizer.state_dict(), cfg.iters, encoder.out_dim)
    frame = render_one(model, encoder, H, W, focal, poses[-1], device, n_samples=cfg.n_samples, near=cfg.near, far=cfg.far)
    imageio.imwrite(f"{cfg.out_dir}/final.png", _to_png(frame))
    print(f"[done] {cfg.iters} iters in {minutes:.2f} min | saved {cfg.ckpt_path} and {cfg.out_dir}/final.png")


if __name__ == "__main__":
    import tyro
    main(tyro.cli(Config))
