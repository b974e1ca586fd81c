"""Training throughput of the three ways to run the reference's training loop on one B200 (VERDICT round 1, items 6 / BASELINE.md
section 5 "second baseline"); BASELINE config 3 shape (4096 rays x 64 samples per step), synthetic 100 x 100 scene:

  fused      this repo's train.py               (engine.Trainer: fused fwd+bwd kernel, gradient scatter, one optimiser launch)
  dropin     the reference's UNMODIFIED train.py (baseline/_ref/src) importing this repo's modules: deferred chain -> one fused
             forward / one fused backward launch per step, torch.optim.Adam + torch.amp.GradScaler as the script asks
  reference  the reference's UNMODIFIED train.py importing the reference's own modules (oracle/_ref/src): its stock PyTorch CUDA
             path (autocast fp16) on the same GPU

Each route runs in a fresh process (previews / checkpoints / logging pushed past the end) with tools/_timing_shims/tqdm in front
of the real tqdm: the unmodified loop `for step in tqdm(range(...))` then reports its own steady-state rate (clock from iteration
--warm to the end of the loop, device synchronised on both sides; start-up and the final render are outside).  Developer / evidence
tool for the GPU box:  python tools/bench_routes.py [--out profiles/r2_routes.json]"""
import argparse
import json
import os
import re
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "tiny-nerf-pytorch_b200")
SHIMS = os.path.join(PKG, "_shims")
TIMING = os.path.join(ROOT, "tools", "_timing_shims")
REF_SCRIPTS = os.path.join(ROOT, "baseline", "_ref", "src")
REF_MODULES = os.path.join(ROOT, "oracle", "_ref", "src")

ap = argparse.ArgumentParser()
ap.add_argument("--out", default=None)
ap.add_argument("--warm", type=int, default=200)
ap.add_argument("--iters", type=int, default=3200)
ap.add_argument("--rays", type=int, default=4096)
ap.add_argument("--samples", type=int, default=64)
args = ap.parse_args()

work = tempfile.mkdtemp(prefix="tnerf_routes_")
os.makedirs(os.path.join(work, "data"))
rng = np.random.default_rng(0)
n_views, H, W = 8, 100, 100
poses = np.tile(np.eye(4, dtype=np.float32), (n_views, 1, 1))
for i in range(n_views):                       # cameras on a circle of radius 4 looking at the origin (content is irrelevant for timing)
    a = 2 * np.pi * i / n_views
    poses[i, :3, 3] = [4 * np.sin(a), 0.0, 4 * np.cos(a)]
    poses[i, :3, :3] = np.array([[np.cos(a), 0, np.sin(a)], [0, 1, 0], [-np.sin(a), 0, np.cos(a)]], dtype=np.float32)
np.savez(os.path.join(work, "data", "tiny_nerf_data.npz"), images=rng.random((n_views, H, W, 3), dtype=np.float32), poses=poses,
         focal=np.float32(138.9))


def run(route, iters):
    far = 10 ** 9
    if route == "fused":
        code = (f"import train; train.main(train.Config(iters={iters}, n_rand={args.rays}, n_samples={args.samples}, log_every={far}, "
                f"preview_every={far}, ckpt_every={far}, resume=False))")
        path = [TIMING, PKG, SHIMS]
    else:
        script = os.path.join(REF_SCRIPTS, "train.py")
        code = (f"import sys, runpy; sys.argv=['train.py','--iters','{iters}','--n-rand','{args.rays}','--n-samples','{args.samples}',"
                f"'--log-every','{far}','--preview-every','{far}','--ckpt-every','{far}','--no-resume'];"
                f"runpy.run_path({script!r}, run_name='__main__')")
        path = [TIMING, PKG, SHIMS] if route == "dropin" else [TIMING, REF_MODULES, SHIMS]
    env = dict(os.environ, PYTHONPATH=os.pathsep.join(path), TNERF_TIMING_WARMUP=str(args.warm))
    for d in ("checkpoints", "outputs"):
        subprocess.run(["rm", "-rf", os.path.join(work, d)])
    t0 = time.time()
    r = subprocess.run([sys.executable, "-c", code], cwd=work, env=env, capture_output=True, text=True, timeout=1500)
    dt = time.time() - t0
    if r.returncode != 0 or f"[done] {iters} iters" not in r.stdout:
        raise RuntimeError(f"{route} ({iters} iters) failed:\n{r.stdout[-1500:]}\n{r.stderr[-2500:]}")
    m = re.search(r"\[timing\] steps=(\d+) seconds=([0-9.]+)", r.stdout)
    if not m:
        raise RuntimeError(f"{route}: no timing line in the output:\n{r.stdout[-1500:]}")
    return int(m.group(1)), float(m.group(2)), dt


res = {"workload": f"train.py loop, {args.rays} rays x {args.samples} samples per step, 100 x 100 synthetic scene, one B200",
       "method": f"{args.iters} iterations, clock from iteration {args.warm} to the end of the loop (device synchronised), fresh process per route",
       "routes": {}}
for route in ("fused", "dropin", "reference"):
    if route != "fused" and not os.path.exists(os.path.join(REF_SCRIPTS, "train.py")):
        res["routes"][route] = {"unavailable": "reference scripts not staged (tools/stage_reference.sh)"}
        continue
    try:
        steps, secs, wall = run(route, args.iters)
        its = steps / secs
        res["routes"][route] = {"it_per_s": its, "ray_samples_per_s": its * args.rays * args.samples, "us_per_step": 1e6 / its,
                                "timed_steps": steps, "process_wall_s": wall}
    except Exception as e:  # noqa: BLE001  (evidence tool: record the failure and go on)
        res["routes"][route] = {"error": str(e)[-1500:]}
    print(route, json.dumps(res["routes"][route]), flush=True)
print(json.dumps(res))
if args.out:
    json.dump(res, open(args.out, "w"), indent=1)
