"""tcgen05.mma issue/execute rate probe (developer tool, GPU box)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tiny-nerf-pytorch_b200"))
import _engine as E
dev = torch.device("cuda:0")
out = torch.zeros(2, dtype=torch.int64, device=dev)
for variant, name in ((0, "TS 1 acc"), (1, "SS 1 acc"), (2, "TS 2 acc"), (3, "TS unroll8")):
    for n in (16, 64, 128, 256):
        for reps in (64, 256):
            for _ in range(2):
                E.check(E.lib().tnerf_umma_rate(n, reps, variant, E.ptr(out), E.stream(dev)))
                torch.cuda.synchronize()
            tot, iss = out.tolist()
            print(f"{name:10s} N={n:3d} reps={reps:3d}: {tot / reps:6.1f} cyc/MMA to completion, {iss / reps:6.1f} cyc/MMA issue  (floor {128 * n / 256:.0f})")
