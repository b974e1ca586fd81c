"""Diagnostic run on the GPU box: UMMA descriptor-convention probe and per-op parity summary.
Writes gpurun_out/probe.log.  Not a test; tests live in tests/."""
import os
import sys
import traceback

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tiny-nerf-pytorch_b200"))
import _engine as E  # noqa: E402
from oracle import oracle as O  # noqa: E402

dev = torch.device("cuda:0")


def selftest(N, K, mode):
    g = torch.Generator().manual_seed(N * 1000 + K + mode)
    m = mode & 3
    a = torch.randn((K, 128) if m == 3 else (128, K), generator=g)
    b = torch.randn((K, N) if m == 2 else (N, K), generator=g)
    d = torch.full((128, N), float("nan"), device=dev)
    a_d, b_d = a.to(dev), b.to(dev)
    E.check(E.lib().tnerf_umma_selftest(E.ptr(a_d), E.ptr(b_d), N, K, mode, E.ptr(d), E.stream(dev)))
    torch.cuda.synchronize()
    ah, bh = a.half().float(), b.half().float()
    A = ah.t() if m == 3 else ah
    B = bh.t() if m == 2 else bh
    ref = A @ B.t()
    return (d.cpu() - ref).abs().max().item(), ref.abs().max().item()


def main():
    print(torch.cuda.get_device_name(0), torch.version.cuda)
    for mode in (0, 1, 2, 3):
        for (N, K) in ((128, 64), (16, 32), (64, 128)):
            try:
                err, mag = selftest(N, K, mode)
                print(f"selftest mode={mode:2d} N={N:3d} K={K:3d}: max err {err:.3e} (ref max {mag:.2f}) {'OK' if err < 2e-2 else 'MISMATCH'}")
            except Exception as ex:  # noqa: BLE001
                print(f"selftest mode={mode} N={N} K={K}: EXC {ex}")
    sys.stdout.flush()


if __name__ == "__main__":
    try:
        main()
    except Exception:  # noqa: BLE001
        traceback.print_exc()
        sys.exit(1)
