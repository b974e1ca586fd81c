"""Parity evidence for DESIGN.md section 7 (run on the GPU box): per test case of tests/test_gpu_fused.py::test_fused_render_vs_oracle
the largest rgb / depth / acc error of the fp16 tensor-core path against the oracle on the rays away from the delta_last
discontinuity, and for the set-aside rays which branch of the discontinuity the engine landed on."""
import json
import math
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tiny-nerf-pytorch_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import oracle as O  # noqa: E402
import engine  # noqa: E402
from encoding import PositionalEncoding  # noqa: E402
from test_gpu_fused import make_model, random_rays  # noqa: E402

dev = torch.device("cuda:0")
cases = [dict(cfg=(63, 128, 4, 2), S=64, n=4096, jitter=True, scale=1.0), dict(cfg=(63, 128, 4, 2), S=64, n=4096, jitter=False, scale=2.0),
         dict(cfg=(39, 128, 4, 2), S=32, n=2048, jitter=True, scale=2.0), dict(cfg=(63, 128, 4, 2), S=192, n=512, jitter=False, scale=2.0),
         dict(cfg=(63, 256, 4, 2), S=192, n=512, jitter=False, scale=1.5), dict(cfg=(63, 256, 4, 2), S=64, n=2048, jitter=True, scale=1.5)]
rows = []
for case in cases:
    ind = case["cfg"][0]
    L = (ind - 3) // 6
    enc = PositionalEncoding(L, True).to(dev)
    model, p = make_model(case["cfg"], 5, dev, case["scale"])
    n, S = case["n"], case["S"]
    ro, rd = random_rays(n, 11)
    u = torch.rand(n, S, generator=torch.Generator().manual_seed(12)) if case["jitter"] else None
    kw = dict(num_freqs=L, include_input=True, depth=case["cfg"][2], skip_at=case["cfg"][3])
    with torch.no_grad():
        comp, depth, acc = engine.render_rays(model, enc, ro.to(dev), rd.to(dev), 2.0, 6.0, S, t_rand=None if u is None else u.to(dev), precision="f16")
    comp, depth, acc = comp.cpu(), depth.cpu(), acc.cpu()
    oc, od, oa, _ = O.render_rays(p, ro, rd, 2.0, 6.0, S, u, **kw)
    fc, fd, fa, _ = O.render_rays_last_flipped(p, ro, rd, 2.0, 6.0, S, u, **kw)
    pre = O.last_sample_sigma_pre(p, ro, rd, 2.0, 6.0, S, u, **kw)
    e_c, e_d, e_a = (comp - oc).abs().amax(1), (depth - od).abs().reshape(-1), (acc - oa).abs().reshape(-1)
    f_c, f_d, f_a = (comp - fc).abs().amax(1), (depth - fd).abs().reshape(-1), (acc - fa).abs().reshape(-1)
    keep = pre.abs() > 4e-3
    ok_as_is = (e_c < 2e-3) & (e_a < 2e-3) & (e_d < 2e-3 * 6)
    ok_flip = (f_c < 2e-3) & (f_a < 2e-3) & (f_d < 2e-3 * 6)
    row = dict(case=str(case), rays=n, kept=int(keep.sum()), max_rgb=float(e_c[keep].max()), max_depth=float(e_d[keep].max()), max_acc=float(e_a[keep].max()),
               p999_depth=float(e_d[keep].quantile(0.999)), depth_over_2e3=int((e_d[keep] > 2e-3).sum()),
               rel_depth=float((e_d[keep] / od.reshape(-1)[keep].clamp_min(1e-3)).max()),
               all_rays_violating_as_is=int((~ok_as_is).sum()), of_which_on_flipped_branch=int((~ok_as_is & ok_flip).sum()),
               unexplained=int((~ok_as_is & ~ok_flip).sum()), unexplained_min_abs_pre=float(pre[~ok_as_is & ~ok_flip].abs().min()) if (~ok_as_is & ~ok_flip).any() else None,
               set_aside=int((~keep).sum()), set_aside_ok_as_is=int((~keep & ok_as_is).sum()), set_aside_flipped=int((~keep & ~ok_as_is & ok_flip).sum()))
    rows.append(row)
    print(json.dumps(row), flush=True)
if len(sys.argv) > 1:
    json.dump(rows, open(sys.argv[1], "w"), indent=1)
