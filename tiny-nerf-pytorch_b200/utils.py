"""utils.mse2psnr -- drop-in for the reference's src/utils.py:14-15 (a scalar; the fused trainer
computes loss and PSNR on device with tnerf_mse_psnr instead)."""
import torch


def mse2psnr(mse: torch.Tensor) -> torch.Tensor:
    """PSNR in dB of a mean-squared error, floored at 1e-10."""
    return torch.log10(torch.clamp_min(mse, 1e-10)) * -10.0
