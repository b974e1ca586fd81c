"""PSNR-parity evidence for DESIGN.md section 7 (GPU box): held-out-view PSNR of the fused engine against the CPU oracle after the
same K steps from the same state, over several seeds and both precisions.  usage: psnr_report.py [K] [out.json]"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tiny-nerf-pytorch_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import make_data  # noqa: E402
from test_gpu_scripts import psnr_parity_run  # noqa: E402

K = int(sys.argv[1]) if len(sys.argv) > 1 else 300
d = make_data.make_scene(n_views=6, H=32, W=32, focal=44.0, n_samples=64)
rows = []
for prec, seeds in (("f16", (7, 8, 9, 10, 11)), ("f32", (7, 8))):
    for seed in seeds:
        r, o = psnr_parity_run(d, prec, seed, K)
        rows.append({"precision": prec, "seed": seed, "steps": K, "psnr_oracle_db": r, "psnr_engine_db": o, "gap_db": abs(r - o)})
        print(json.dumps(rows[-1]), flush=True)
if len(sys.argv) > 2:
    json.dump(rows, open(sys.argv[2], "w"), indent=1)
