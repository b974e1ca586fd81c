"""tcgen05.mma issue / completion rate probe (developer tool, run on the GPU box)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tiny-nerf-pytorch_b200"))
import _engine as E  # noqa: E402

dev = torch.device("cuda:0")
out = torch.zeros(8, dtype=torch.int64, device=dev)
CASES = (("SS 1 warp", 8), ("SS 2 warps", 9), ("TS 1 warp", 8 + 16), ("TS 2 warps", 9 + 16), ("TS 2 warps, B streamed", 9 + 16 * 5),
         ("TS 2 warps, B streamed, +traffic", 9 + 16 * 7), ("SS 2 warps, B streamed, +traffic", 9 + 16 * 6))
for name, variant in CASES:
    for n in (16, 64, 128, 256):
        for reps in (512,):
            for _ in range(2):
                out.zero_()
                E.check(E.lib().tnerf_umma_rate(n, reps, variant, E.ptr(out), E.stream(dev)))
                torch.cuda.synchronize()
            o = out.tolist()
            nw = ((variant - 8) & 3) + 1 if variant >= 8 else 1
            tot = max(o[0:2 * nw:2]); iss = max(o[1:2 * nw:2])
            print(f"{name:34s} N={n:3d} reps={reps:3d}x{nw}: {tot / (reps * nw):6.1f} cyc/MMA to completion, {iss / reps:6.1f} cyc/MMA issue per warp (floor {128 * n / 256:.0f})")
