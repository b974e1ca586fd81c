"""Stress loop for the fused render kernel: many launches over several (n_samples, n_rays) shapes incl. ragged ray counts and
multi-tile units (192 samples), progress per batch (run under `timeout`).  Developer tool."""
import ctypes as C
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tiny-nerf-pytorch_b200"))
import _engine as E  # noqa: E402
import engine  # noqa: E402
from nerf import TinyNeRF  # noqa: E402

dev = torch.device("cuda:0")
torch.manual_seed(0)
hidden = int(os.environ.get("STRESS_HIDDEN", "128"))          # 256 = the CTA-pair kernel of BASELINE config 4
model = TinyNeRF(63, hidden, 4, 2).to(dev)
h = E.handle_for(model, dev); h.set_encoding(10, True); h.ensure_packed(force=True)
pose = torch.eye(4, device=dev); pose[2, 3] = 4.0
cases = [(int(a.split("x")[0]), int(a.split("x")[1])) for a in sys.argv[1].split(",")]
launches = int(sys.argv[2]) if len(sys.argv) > 2 else 200
for S, n in cases:
    comp, depth, acc = torch.empty(n, 3, device=dev), torch.empty(n, 1, device=dev), torch.empty(n, 1, device=dev)
    jit = torch.rand(n, S, device=dev)
    rs = engine.ray_source(c2w=pose, H=1000, W=1000, focal=1111.1, first_ray=0)
    t0 = time.time()
    for k in range(launches):
        E.check(E.lib().tnerf_render_fwd(h.h, C.byref(rs), n, 2.0, 6.0, S, E.ptr(jit) if k & 1 else None, 1, 0, E.ptr(comp), E.ptr(depth), E.ptr(acc),
                                         None, None, E.stream(dev)))
        if k % 50 == 49:
            torch.cuda.synchronize()
    torch.cuda.synchronize()
    print(f"S={S} n={n}: {launches} launches ok, {time.time() - t0:.1f} s, finite={bool(torch.isfinite(comp).all())}", flush=True)
print("STRESS OK", flush=True)
