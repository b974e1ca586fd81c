"""Phase timeline of the two-stream training kernel (CTA 0): clock64 stamps of the stream-0 drain warp, the stream-0
MMA issuer, the stream-0 sample warp and the stream-1 drain warp, printed as cycle deltas.  Developer tool (GPU box)."""
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
pix = torch.randint(0, 10000, (n,), device=dev)
target = torch.rand(n, 3, device=dev)
u = torch.rand(n, S, device=dev)
dbg = torch.zeros(2048, dtype=torch.int64, device=dev)
h = E.handle_for(model, dev)
for it in range(3):
    dbg.zero_()
    E.check(E.lib().tnerf_set_debug_buffer(h.h, E.ptr(dbg) if it == 2 else None))
    tr.step_pixels(pose, 100, 100, 138.9, pix, target, u)
torch.cuda.synchronize()
d = dbg.cpu().tolist()
t0 = min(x for x in d if x)
for name, off in (("drain WG0", 0), ("issuer 0", 256), ("sample warp 0", 512), ("drain WG1", 768)):
    xs = [x for x in d[off:off + 256] if x]
    print(f"--- {name}: {len(xs)} stamps; (cycles since kernel start) pairs = wait begin -> wait end")
    line = []
    for i, x in enumerate(xs[:int(sys.argv[1]) if len(sys.argv) > 1 else 70]):
        line.append(f"{x - t0:7d}(+{x - xs[i - 1] if i else 0:5d})")
        if len(line) == 6:
            print("  " + " ".join(line)); line = []
    if line:
        print("  " + " ".join(line))

for name, off in (("drain WG0", 0), ("drain WG1", 768)):
    print(f"{name}: loop end {d[off + 253] - t0}, all GEMMs done {d[off + 254] - t0}, slab flushed {d[off + 255] - t0}")
ns = d[252] - d[251]
cyc = d[255] - d[0]
print(f"drain WG0 thread 0 lifetime: {cyc} cycles in {ns} ns -> SM clock {cyc / ns * 1e3:.0f} MHz")
