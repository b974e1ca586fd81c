"""Developer check of the hidden=256 CTA-pair render kernel: error pattern against the CPU oracle on a few rays, then the
time of a BASELINE-config-4 slice (rows of an 800x800 frame, 192 samples/ray).  Usage: python tools/wide_check.py [rows]"""
import math
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tiny-nerf-pytorch_b200"))
import engine  # noqa: E402
from encoding import PositionalEncoding  # noqa: E402
from nerf import TinyNeRF  # noqa: E402
from oracle import oracle as O  # noqa: E402

dev = torch.device("cuda:0")
p = O.init_params(63, 256, 4, 2, seed=5)
p = {k: (v * 1.5 if k.endswith("weight") else v) for k, v in p.items()}
p["sigma.0.bias"] = p["sigma.0.bias"] + 0.3
model = TinyNeRF(63, 256, 4, 2)
model.load_state_dict(p)
model = model.to(dev)
enc = PositionalEncoding(10, True).to(dev)

import _engine as E  # noqa: E402
dbg = None
if os.environ.get("WIDE_DBG"):
    dbg = torch.zeros(2048, dtype=torch.int64, device=dev)
    E.check(E.lib().tnerf_set_debug_buffer(E.handle_for(model, dev).h, E.ptr(dbg)))


def show_dbg(tag):
    if dbg is None:
        return
    torch.cuda.synchronize()
    d = dbg.cpu().tolist()
    for b in range(16):
        for w in range(9):
            site, par = d[512 + (b * 9 + w) * 2], d[512 + (b * 9 + w) * 2 + 1]
            if site:
                print(f"   [{tag}] CTA {b} warp {w}: first timed-out wait = site {site} (parity {par})", flush=True)
    dbg.zero_()


for S, n in ((32, 8), (64, 300), (192, 301)):
    g = torch.Generator().manual_seed(11)
    pose = O.look_at_pose(1.0, 0.5)
    ro, rd = O.get_rays(64, 64, 80.0, pose)
    idx = torch.randint(0, 64 * 64, (n,), generator=g)
    ro, rd = ro[idx].contiguous(), rd[idx].contiguous()
    u = torch.rand(n, S, generator=g)
    with torch.no_grad():
        comp, depth, acc = engine.render_rays(model, enc, ro.to(dev), rd.to(dev), 2.0, 6.0, S, t_rand=u.to(dev), precision="f16")
        c32, d32, a32 = engine.render_rays(model, enc, ro.to(dev), rd.to(dev), 2.0, 6.0, S, t_rand=u.to(dev), precision="f32")
    torch.cuda.synchronize()
    show_dbg(f"S={S}")
    oc, od, oa, _ = O.render_rays(p, ro, rd, 2.0, 6.0, S, u)
    e = (comp.cpu() - oc).abs().max(dim=1).values
    print(f"S={S} n={n}: f16 max err rgb {e.max():.2e} (median {e.median():.2e}) acc {(acc.cpu() - oa).abs().max():.2e} "
          f"depth {(depth.cpu() - od).abs().max():.2e} | f32 path rgb {(c32.cpu() - oc).abs().max():.2e}", flush=True)
    bad = (e > 2e-3).nonzero().flatten().tolist()
    if bad:
        print("   rays over 2e-3:", bad[:40], "..." if len(bad) > 40 else "", flush=True)

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 100
H = W = 800
pose = O.look_at_pose(0.7, 0.5).to(dev)
n = rows * W
import ctypes as C  # noqa: E402
h = E.handle_for(model, dev)
h.set_encoding(10, True)
for prec, name in ((engine._PREC["f16"], "f16 tcgen05 pair kernel"), (engine._PREC["f32"], "f32 path")):
    if prec == engine._PREC["f16"]:
        h.ensure_packed(force=True)
    comp = torch.empty(n, 3, device=dev); depth = torch.empty(n, 1, device=dev); acc = torch.empty(n, 1, device=dev)
    rs = engine.ray_source(c2w=pose, H=H, W=W, focal=1111.11, first_ray=300 * W)
    reps = 3 if prec == engine._PREC["f16"] else 1
    ts = []
    for _ in range(reps + 1):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        E.check(E.lib().tnerf_render_fwd(h.h, C.byref(rs), n, 2.0, 6.0, 192, None, 1, prec, E.ptr(comp), E.ptr(depth), E.ptr(acc), None, None,
                                         E.stream(dev)))
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ms = min(ts[1:]) if len(ts) > 1 else ts[0]
    flop = n * 192 * 459776.0
    print(f"{name}: {rows} rows x 800 x 192 samples: {ms:.2f} ms = {n / ms * 1e-3:.2f} M rays/s, {flop / ms * 1e-9:.0f} TFLOP/s "
          f"({flop / ms * 1e-9 / 1630 * 100:.0f}% of 1630)", flush=True)
    if prec == engine._PREC["f16"]:
        ref16 = comp.clone()
    else:
        print(f"   f16 vs f32 path on the slice: max |d rgb| {(ref16 - comp).abs().max():.2e}", flush=True)
