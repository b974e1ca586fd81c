"""Pose-batched rendering (tnerf_render_frames) against one launch per pose: a 60-pose spiral of 100x100x64 frames (make_gif.py's
workload, src/make_gif.py:22-27).  Evidence tool, run on the GPU box."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tiny-nerf-pytorch_b200"))
import engine  # noqa: E402
from camera import spiral_poses  # noqa: E402
from encoding import PositionalEncoding  # noqa: E402
from nerf import TinyNeRF  # noqa: E402
from train import render_one  # noqa: E402

dev = torch.device("cuda:0")
torch.manual_seed(0)
enc = PositionalEncoding(10, True).to(dev)
model = TinyNeRF(63, 128, 4, 2).to(dev)
ref = torch.eye(4, device=dev); ref[2, 3] = 4.0
path = spiral_poses(ref, n_frames=60, radius=0.3)
H = W = 100
focal = 138.9


def timed(fn, reps=10):
    fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


t_batch = timed(lambda: engine.render_frames(model, enc, H, W, focal, path, n_samples=64))
t_loop = timed(lambda: [render_one(model, enc, H, W, focal, path[i], dev, n_samples=64) for i in range(60)])
rays = 60 * H * W
print(f"60 poses x {H}x{W}x64: one launch {t_batch:.3f} ms ({rays / t_batch * 1e-3:.1f} M rays/s, {t_batch / 60 * 1e3:.1f} us per frame); "
      f"one render_one per pose {t_loop:.3f} ms ({rays / t_loop * 1e-3:.1f} M rays/s, {t_loop / 60 * 1e3:.1f} us per frame)")
