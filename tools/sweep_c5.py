"""BASELINE config 5: ray-batch sweep 2^14 .. 2^22 rays x 128 samples, fused fwd+bwd (tnerf_train_fwd_bwd, no optimiser), reporting
ray-samples/s and the fraction of the measured tensor peak per batch size.  Inputs (pixel ids, targets, explicit jitter) are
device resident; from 2^17 rays on the jitter tensor alone exceeds the L2.  Evidence tool, run on the GPU box."""
import argparse
import ctypes as C
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tiny-nerf-pytorch_b200"))
import _engine as E  # noqa: E402
import engine  # noqa: E402
from encoding import PositionalEncoding  # noqa: E402
from nerf import TinyNeRF  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--out", default=None)
ap.add_argument("--max-log2", type=int, default=22)
ap.add_argument("--min-log2", type=int, default=14)
ap.add_argument("--samples", type=int, default=128)
args = ap.parse_args()
dev = torch.device("cuda:0")
torch.manual_seed(0)
enc = PositionalEncoding(10, True).to(dev)
model = TinyNeRF(63, 128, 4, 2).to(dev)
tr = engine.Trainer(model, enc, n_samples=args.samples)
pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
peak = json.load(open(pk))["bf16_tflops"] if os.path.exists(pk) else 1590.0
FLOP = 362496
pose = torch.eye(4, device=dev); pose[2, 3] = 4.0
rows = []
S = args.samples
for lg in range(args.min_log2, args.max_log2 + 1):
    n = 1 << lg
    pix = torch.randint(0, 800 * 800, (n,), device=dev)
    tgt = torch.rand(n, 3, device=dev)
    jit = torch.rand(n, S, device=dev)
    rs = engine.ray_source(c2w=pose, H=800, W=800, focal=1111.1, pixel_index=pix)

    def call():
        E.check(E.lib().tnerf_train_fwd_bwd(tr.h.h, C.byref(rs), E.ptr(tgt), n, 2.0, 6.0, S, E.ptr(jit), 1, tr.prec, 3.0 * n, None,
                                            E.ptr(tr.loss_view), E.ptr(tr.gbuf), E.stream(dev)))
    reps = max(3, min(20, (1 << 24) // n))
    for _ in range(2):
        call()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        call()
    b.record()
    torch.cuda.synchronize()
    tr.gbuf.zero_()
    sec = a.elapsed_time(b) / reps * 1e-3
    r = {"rays": n, "samples_per_ray": S, "seconds": sec, "ray_samples_per_s": n * S / sec, "tflops": n * S * FLOP / sec / 1e12,
         "frac_of_measured_peak": n * S * FLOP / sec / 1e12 / peak, "input_mb": n * (S * 4 + 20) / 1e6}
    rows.append(r)
    print(f"2^{lg:2d} rays x {S}: {sec * 1e3:9.3f} ms  {r['ray_samples_per_s'] / 1e9:6.3f} G ray-samples/s  {r['tflops']:7.1f} TFLOP/s  {100 * r['frac_of_measured_peak']:5.1f} % of {peak:.0f}", flush=True)
    del pix, tgt, jit
out = {"config": "C5 ray-batch sweep, fwd+bwd, L=10 hidden=128", "peak_tflops": peak, "flop_per_sample": FLOP, "rows": rows}
if args.out:
    json.dump(out, open(args.out, "w"), indent=1)
