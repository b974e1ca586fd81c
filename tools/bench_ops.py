"""Achieved HBM bandwidth of the stand-alone drop-in kernels (the fused kernels are tensor-bound, the HBM roofline
applies to get_rays / stratified_samples / PositionalEncoding / volume_render and to the gradient plumbing).  Algorithmic bytes per
unit as stated in DESIGN.md section 5; CUDA events on the launching stream; the working set of every case exceeds the 126 MB L2.
Writes one JSON object (also to profiles/ when --out is given).  Developer / evidence tool, run on the GPU box."""
import argparse
import ctypes as C
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tiny-nerf-pytorch_b200"))
import _engine as E  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--out", default=None)
args = ap.parse_args()
dev = torch.device("cuda:0")
peak = 6552.6
pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
if os.path.exists(pk):
    peak = json.load(open(pk))["hbm_gbs"]
lib, st = E.lib(), E.stream(dev)


def timed(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e-3


res = {}
S, L, D = 64, 10, 63
# get_rays: 2048 x 2048 frame -> 24 B/ray written (4.2 M rays, 100 MB) x 2 frames alternating
Hh = Ww = 2048
n = Hh * Ww
pose = torch.eye(4, device=dev); pose[2, 3] = 4.0
ro = [torch.empty(n, 3, device=dev) for _ in range(2)]
rd = [torch.empty(n, 3, device=dev) for _ in range(2)]
k = [0]


def f_rays():
    i = k[0] & 1; k[0] += 1
    E.check(lib.tnerf_get_rays(Hh, Ww, 1111.0, E.ptr(pose), 0, n, E.ptr(ro[i]), E.ptr(rd[i]), st))


t = timed(f_rays)
res["get_rays"] = {"units": n, "bytes_per_unit": 24, "seconds": t}


# reference for a WRITE-ONLY stream of the same size and launch pattern: torch's fill of the same two alternating 100 MB pairs
# (the roofline denominator is a copy, read + write; a pure write stream does not reach it)
def f_fill():
    i = k[0] & 1; k[0] += 1
    ro[i].fill_(1.0); rd[i].fill_(2.0)


t = timed(f_fill)
res["write_only_reference_fill"] = {"units": n, "bytes_per_unit": 24, "seconds": t}
# stratified: n2 rays x 64 samples: in 24 + 4S (jitter), out 16 S (z + pts)
n2 = 1 << 18
jit = torch.rand(n2, S, device=dev)
z = torch.empty(n2, S, device=dev); pts = torch.empty(n2, S, 3, device=dev)
o2, d2 = torch.randn(n2, 3, device=dev), torch.randn(n2, 3, device=dev)


def f_strat():
    E.check(lib.tnerf_stratified(E.ptr(o2), 3, E.ptr(d2), n2, S, 2.0, 6.0, None, None, E.ptr(jit), E.ptr(z), E.ptr(pts), st))


t = timed(f_strat)
res["stratified_samples"] = {"units": n2, "bytes_per_unit": 24 + 4 * S + 16 * S, "seconds": t}
# positional encoding: 2^22 points: 12 B in + 4 D out
npts = 1 << 22
x = torch.randn(npts, 3, device=dev); enc = torch.empty(npts, D, device=dev)


def f_enc():
    E.check(lib.tnerf_posenc(E.ptr(x), npts, L, 1, E.ptr(enc), st))


t = timed(f_enc)
res["positional_encoding"] = {"units": npts, "bytes_per_unit": 12 + 4 * D, "seconds": t}
# volume_render forward: n2 rays x 64: (20 S + 12) in, 20 + 4 S out
rgb = torch.rand(n2, S, 3, device=dev); sig = torch.rand(n2, S, device=dev)
comp = torch.empty(n2, 3, device=dev); dep = torch.empty(n2, 1, device=dev); acc = torch.empty(n2, 1, device=dev); w = torch.empty(n2, S, device=dev)


def f_comp():
    E.check(lib.tnerf_composite_fwd(E.ptr(rgb), E.ptr(sig), E.ptr(z), S, E.ptr(d2), n2, S, 1, E.ptr(comp), E.ptr(dep), E.ptr(acc), E.ptr(w), st))


t = timed(f_comp)
res["volume_render_fwd"] = {"units": n2, "bytes_per_unit": 20 * S + 12 + 20 + 4 * S, "seconds": t}
# volume_render backward: reads rgb, sigma, z, d, gC (+ gW), writes g_rgb, g_sigma: (20 S + 12 + 12 + 4 S) in, 16 S out
gC = torch.rand(n2, 3, device=dev); gW = torch.rand(n2, S, device=dev)
g_rgb = torch.empty_like(rgb); g_sig = torch.empty_like(sig)


def f_compb():
    E.check(lib.tnerf_composite_bwd(E.ptr(rgb), E.ptr(sig), E.ptr(z), S, E.ptr(d2), n2, S, 1, E.ptr(gC), None, None, E.ptr(gW), E.ptr(g_rgb), E.ptr(g_sig), st))


t = timed(f_compb)
res["volume_render_bwd"] = {"units": n2, "bytes_per_unit": 20 * S + 24 + 4 * S + 16 * S, "seconds": t}
for name, r in res.items():
    r["achieved_gbs"] = r["units"] * r["bytes_per_unit"] / r["seconds"] / 1e9
    r["frac_of_measured_hbm_peak"] = r["achieved_gbs"] / peak
    r["working_set_mb"] = r["units"] * r["bytes_per_unit"] / 1e6
out = {"peak_hbm_gbs": peak, "peak_source": "MEASURED_PEAKS.json" if os.path.exists(pk) else "fallback", "kernels": res}
print(json.dumps(out, indent=1))
if args.out:
    json.dump(out, open(args.out, "w"), indent=1)
