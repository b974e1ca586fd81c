"""Stress loop for the fused training kernel: many long launches (hundreds of tiles per stream) at several n_samples, with a
progress line per batch of launches so that a hang can be localised (run under `timeout`).  Developer tool."""
import ctypes as C
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tiny-nerf-pytorch_b200"))
import _engine as E  # noqa: E402
import engine  # noqa: E402
from encoding import PositionalEncoding  # noqa: E402
from nerf import TinyNeRF  # noqa: E402

dev = torch.device("cuda:0")
torch.manual_seed(0)
enc = PositionalEncoding(10, True).to(dev)
model = TinyNeRF(63, 128, 4, 2).to(dev)
pose = torch.eye(4, device=dev); pose[2, 3] = 4.0
cases = [(int(a.split("x")[0]), int(a.split("x")[1])) for a in sys.argv[1].split(",")]     # e.g. 128x131072,64x262144
launches = int(sys.argv[2]) if len(sys.argv) > 2 else 200
for S, n in cases:
    tr = engine.Trainer(model, enc, n_samples=S)
    pix = torch.randint(0, 640000, (n,), device=dev); tgt = torch.rand(n, 3, device=dev); jit = torch.rand(n, S, device=dev)
    rs = engine.ray_source(c2w=pose, H=800, W=800, focal=1111.1, pixel_index=pix)
    t0 = time.time()
    for k in range(launches):
        E.check(E.lib().tnerf_train_fwd_bwd(tr.h.h, C.byref(rs), E.ptr(tgt), n, 2.0, 6.0, S, E.ptr(jit), 1, tr.prec, 3.0 * n, None,
                                            E.ptr(tr.loss_view), E.ptr(tr.gbuf), None, None, E.stream(dev)))
        if k % 10 == 9:
            torch.cuda.synchronize()
            print(f"S={S} n={n}: {k + 1} launches ok, {time.time() - t0:.1f} s", flush=True)
            tr.gbuf.zero_()
print("STRESS OK", flush=True)
