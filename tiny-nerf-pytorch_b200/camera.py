"""camera.spiral_poses -- same contract as the reference's src/camera.py:4-12 (host-side 4x4 math)."""
import math

import torch


def spiral_poses(c2w_ref: torch.Tensor, n_frames: int = 60, radius: float = 0.3) -> torch.Tensor:
    """(n_frames,4,4): the reference pose translated along a circle of ``radius`` in its own x/y plane."""
    ang = torch.linspace(0, 2 * math.pi, n_frames, device=c2w_ref.device)
    shift = torch.eye(4, device=c2w_ref.device, dtype=c2w_ref.dtype).repeat(n_frames, 1, 1)
    shift[:, 0, 3] = (radius * torch.cos(ang)).to(c2w_ref.dtype)
    shift[:, 1, 3] = (radius * torch.sin(ang)).to(c2w_ref.dtype)
    return c2w_ref.unsqueeze(0) @ shift
