"""Phase timeline of the fused forward kernel (CTA 0): clock64 stamps of warpgroup 0's row 0 and of the MMA
issuer, printed as cycle deltas.  Developer tool (run on the GPU box)."""
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
h = E.handle_for(model, dev); h.set_encoding(10, True); h.ensure_packed(force=True)
n, S = 10000, 64
pose = torch.eye(4, device=dev); pose[2, 3] = 4.0
comp = torch.empty(n, 3, device=dev)
dbg = torch.zeros(1024, dtype=torch.int64, device=dev)
rs = engine.ray_source(c2w=pose, H=100, W=100, focal=138.9, first_ray=0)
for it in range(3):
    dbg.zero_()
    E.check(E.lib().tnerf_set_debug_buffer(h.h, E.ptr(dbg) if it == 2 else None))
    E.check(E.lib().tnerf_render_fwd(h.h, C.byref(rs), n, 2.0, 6.0, S, None, 1, 0, E.ptr(comp), None, None, None, None, E.stream(dev)))
torch.cuda.synchronize()
raw = dbg.cpu().tolist()
d = [x for x in raw[:256] if x]
names = ["loop top", "xfree(t) + features(t+1) stored", "features(t+2) in registers", "heads(t) ready", "composite(t) done"]
t0 = d[0]
print("sample warpgroup 0, thread 0 (cycles since start, delta):")
for i, x in enumerate(d[:6 * len(names)]):
    print(f"  {names[i % len(names)]:34s} {x - t0:8d}  +{(x - d[i - 1]) if i else 0}")

for name, off in (("epilogue warpgroup 0 (wait begin, wait end)", 256), ("issuer 0 (wait begin, wait end)", 512)):
    xs = [x for x in raw[off:off + 250] if x]
    print(name)
    line = []
    for i, x in enumerate(xs[0:120]):
        line.append(f"{x - t0:7d}(+{x - xs[i - 1] if i else 0:5d})")
        if len(line) == 6:
            print("  " + " ".join(line)); line = []
