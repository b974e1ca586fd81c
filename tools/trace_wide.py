"""Phase timeline of the hidden=256 CTA-pair render kernel (CTA 0): clock64 stamps of the epilogue warp 0 (before / after each
accumulator wait) and of sample warp 4, printed as cycle deltas.  Developer tool (run on the GPU box)."""
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
model = TinyNeRF(63, 256, 4, 2).to(dev)
h = E.handle_for(model, dev); h.set_encoding(10, True); h.ensure_packed(force=True)
S = int(sys.argv[1]) if len(sys.argv) > 1 else 192
n = 148 * 8 * 128 // S * 2
pose = torch.eye(4, device=dev); pose[2, 3] = 4.0
comp = torch.empty(n, 3, device=dev)
dbg = torch.zeros(2048, dtype=torch.int64, device=dev)
rs = engine.ray_source(c2w=pose, H=800, W=800, focal=1111.1, first_ray=0)
for it in range(3):
    dbg.zero_()
    E.check(E.lib().tnerf_set_debug_buffer(h.h, E.ptr(dbg) if it == 2 else None))
    E.check(E.lib().tnerf_render_fwd(h.h, C.byref(rs), n, 2.0, 6.0, S, None, 1, 0, E.ptr(comp), None, None, None, None, E.stream(dev)))
torch.cuda.synchronize()
raw = dbg.cpu().tolist()
d = [x for x in raw[:250] if x]
names = ["loop top", "xfree(t) + features(t+1) stored", "features(t+2) in registers", "heads(t) ready", "composite(t) done"]
t0 = d[0]
print("sample warp 4, lane 0 (cycles since start, delta):")
for i, x in enumerate(d[:4 * len(names)]):
    print(f"  {names[i % len(names)]:34s} {x - t0:8d}  +{(x - d[i - 1]) if i else 0}")
xs = [x for x in raw[256:256 + 250] if x]
print("epilogue warp 0: per quarter (wait begin -> wait end = idle, wait end -> next wait begin = busy)")
for t in range(min(3, len(xs) // 32)):
    row = xs[32 * t:32 * t + 33]
    idle = [row[2 * i + 1] - row[2 * i] for i in range(16)]
    busy = [row[2 * i + 2] - row[2 * i + 1] for i in range(16) if 2 * i + 2 < len(row)]
    print(f"  tile {t}: start {row[0] - t0}; idle {idle}")
    print(f"          busy {busy}; tile total {row[-1] - row[0] if len(row) == 33 else 0}")
