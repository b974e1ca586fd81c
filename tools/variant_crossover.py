"""Rolled vs unrolled tile program of the training kernel (TNERF_TRAIN_UNROLL_FROM) over batch sizes: where does the unrolled variant
start to win?  Developer tool (run on the GPU box)."""
import ctypes as C
import os
import sys

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
tr = engine.Trainer(model, enc, n_samples=int(sys.argv[1]) if len(sys.argv) > 1 else 64)
pose = torch.eye(4, device=dev); pose[2, 3] = 4.0
S = int(sys.argv[1]) if len(sys.argv) > 1 else 64
for n in (100, 1000, 2048, 4096, 4097, 6144, 8192, 16384, 65536):
    nsets = max(2, int(140e6 / (n * (S * 4 + 20))) + 1)
    pix = torch.randint(0, 10000, (nsets, n), device=dev); tgt = torch.rand(nsets, n, 3, device=dev); jit = torch.rand(nsets, n, S, device=dev)
    rss = [engine.ray_source(c2w=pose, H=100, W=100, focal=138.9, pixel_index=pix[k]) for k in range(nsets)]
    res = []
    for mode, sync in (("1000000000", "0"), ("1", "1")):
        tr.h.set_option("unroll_from", int(mode))
        tr.h.set_option("train_sync", int(sync))
        reps = max(20, 400000 // n)
        for timed in (False, True):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for i in range(reps):
                k = i % nsets
                E.check(E.lib().tnerf_train_fwd_bwd(tr.h.h, C.byref(rss[k]), E.ptr(tgt[k]), n, 2.0, 6.0, S, E.ptr(jit[k]), 1, tr.prec, 3.0 * n, None,
                                                    E.ptr(tr.loss_view), E.ptr(tr.gbuf), None, None, E.stream(dev)))
            b.record()
            torch.cuda.synchronize()
        res.append(a.elapsed_time(b) / reps * 1e3)
    tiles = n * S / 64 / 296
    print(f"n={n:6d} ({tiles:6.1f} tiles/stream): rolled / half a tile apart {res[0]:8.1f} us, unrolled / in phase {res[1]:8.1f} us  ({res[0] / res[1]:.3f}x)", flush=True)

