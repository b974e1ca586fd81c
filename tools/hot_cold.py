"""Training-kernel time with L2-hot inputs (one input set) vs L2-cold inputs (124 rotating sets = 140 MB), plus the SM clock the
kernel saw (cycles / globaltimer of CTA 0).  Developer tool."""
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
tr = engine.Trainer(model, enc, n_samples=64)
n, S = 4096, 64
pose = torch.eye(4, device=dev); pose[2, 3] = 4.0
for nsets in (1, 124):
    pix = torch.randint(0, 10000, (nsets, n), device=dev); tgt = torch.rand(nsets, n, 3, device=dev); jit = torch.rand(nsets, n, S, device=dev)
    rss = [engine.ray_source(c2w=pose, H=100, W=100, focal=138.9, pixel_index=pix[k]) for k in range(nsets)]
    dbg = torch.zeros(2048, dtype=torch.int64, device=dev)
    E.check(E.lib().tnerf_set_debug_buffer(tr.h.h, E.ptr(dbg)))
    for reps in (5, 100):
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for i in range(reps):
            k = i % nsets
            E.check(E.lib().tnerf_train_fwd_bwd(tr.h.h, C.byref(rss[k]), E.ptr(tgt[k]), n, 2.0, 6.0, S, E.ptr(jit[k]), 1, tr.prec, 3.0 * n, None,
                                                E.ptr(tr.loss_view), E.ptr(tr.gbuf), None, None, E.stream(dev)))
        b.record()
        torch.cuda.synchronize()
        d = dbg.cpu().tolist()
        cyc, ns = d[255] - d[0], d[252] - d[251]
        st = [d[1024 + 4 * b] for b in range(148)]; en = [d[1027 + 4 * b] for b in range(148)]
        le = [d[1025 + 4 * b] for b in range(148)]; sy = [d[1026 + 4 * b] for b in range(148)]
        life = sorted(e - s0 for s0, e in zip(st, en))
        loop = sorted(x - s0 for s0, x in zip(st, le)); wait = sorted(y - x for x, y in zip(le, sy)); flush = sorted(e - y for y, e in zip(sy, en))
        order = sorted(range(148), key=lambda b: le[b] - st[b])
        print(f"    stream-0 drain loop done after min/med/max {loop[0]}/{loop[74]}/{loop[-1]} ns (fastest CTAs {order[:5]}, slowest {order[-5:]}); "
              f"wait for the other stream {wait[0]}/{wait[74]}/{wait[-1]}; flush {flush[0]}/{flush[74]}/{flush[-1]}", flush=True)
        print(f"    last launch: first CTA start -> last CTA end {max(en) - min(st)} ns; CTA lifetimes min/median/max {life[0]}/{life[74]}/{life[-1]} ns; "
              f"start skew {max(st) - min(st)} ns", flush=True)
        print(f"input sets {nsets:3d}, {reps:3d} back-to-back calls: {a.elapsed_time(b) / reps * 1e3:7.1f} us per call (kernel + slab reduce); "
              f"last launch CTA 0: {cyc} cycles in {ns} ns = {cyc / max(ns, 1) * 1e3:.0f} MHz", flush=True)
    E.check(E.lib().tnerf_set_debug_buffer(tr.h.h, None))
    tr.gbuf.zero_()
