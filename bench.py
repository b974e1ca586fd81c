#!/usr/bin/env python
"""bench.py -- headline measurement of the TinyNeRF ray engine (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload train|render|c1|c4|c5] [--impl reference]

Workloads (BASELINE.json `configs`):
  train  (default) C3: train.py's random-ray batch, 4096 rays x 64 samples per GPU, fused fwd+bwd + Adam (L=10, hidden 128,
         depth 4, skip 2), ray-sharded data parallel (weak scaling), one exchange of the 265 KB gradient per step.
  render C2: full 100x100 view, 64 samples, fused forward; rays/s.
  c1     C1: the tiny_nerf_min.py step (src/tiny_nerf_min.py:1149-1374: 2048 rays x 64 samples fwd+bwd+Adam) and its
         render_image (:1379-1460, 100x100 frame), at L=10 (as coded) and L=6 (as BASELINE.json words it).
  c4     C4: 800x800 frame, 192 samples, hidden 256, rows sharded over the ranks (strong scaling); rays/s.
  c5     C5: ray-batch sweep 2^14..2^22 rays x 128 samples, fused fwd+bwd (no optimiser): one line for --rays (default 2^17),
         `--sweep` adds every batch size to the same line.
One JSON line is printed by rank 0.

Timing protocol (every GPU number): W warm-up steps, then R rounds; a round = barrier + synchronize, two untimed steps (after a
barrier the ranks are milliseconds apart; the exchange kernel's rendezvous re-aligns them), CUDA event, EXACTLY K steps, CUDA
event, barrier + synchronize.  The reported time is the MEDIAN round (max over ranks per round), so `ms_per_step * steps` is the
duration of one timed K-step region.  NVML clocks are sampled from before the first round to after the last.

`--impl reference` times the reference's own CPU implementation of the same step on the host cores: the UNMODIFIED reference
modules staged under oracle/_ref/src (tools/stage_reference.sh; cpu_baseline.kind = "reference"), else the oracle port
(oracle/oracle.py; kind = "port"), with the same config / metric.
"""
import argparse
import importlib.util
import json
import math
import os
import subprocess
import sys
import threading
import time
import types

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "tiny-nerf-pytorch_b200")
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

FOCAL = 138.88888549804688


def flop_per_sample(L=10, hidden=128, bwd=True):
    """2*MAC of the Linear layers: forward; + weight gradients (same) + input gradients of layers 1..3 and
    the heads (the encoding needs none).  L=10, hidden 128: 131 584 / 362 496; hidden 256: 459 776 forward."""
    D = 6 * L + 3
    fwd = 2 * (D * hidden + hidden * hidden + (hidden + D) * hidden + hidden * hidden + hidden * 4)
    dgrad = 2 * (3 * hidden * hidden + hidden * 4)
    return fwd + (fwd + dgrad if bwd else 0)


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return {"hbm_gbs": d["hbm_gbs"], "tf_burst": d["bf16_tflops"], "tf_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                "source": "measured"}
    return {"hbm_gbs": 6650.0, "tf_burst": 1590.0, "tf_sustained": 1400.0, "source": "fallback"}


class ClockSampler(threading.Thread):
    """SM clock and throttle reasons sampled DURING the timed region (B200_PROFILING.md), in-process through NVML
    (nvidia-ml-py) every ~2 ms; falls back to polling nvidia-smi.  Constructed and started BEFORE the first barrier of the timed
    rounds (NVML initialisation takes milliseconds and differs between ranks)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.sm, self.reason_bits, self.stop_flag, self.max_mhz, self.how = index, [], 0, False, None, "nvml"
        self.nv = None
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = index
            if vis and not vis.startswith("GPU"):
                idx = int(vis.split(",")[index])
            self.h = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.nv = pynvml
        except Exception:  # noqa: BLE001
            self.how = "nvidia-smi"

    def start(self):
        self.running = threading.Event()
        super().start()
        self.running.wait(2.0)

    def run(self):
        while not self.stop_flag:
            if self.nv is not None:
                try:
                    self.sm.append(float(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM)))
                    self.reason_bits |= int(self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                except Exception:  # noqa: BLE001
                    pass
                self.running.set()
                time.sleep(0.002)
            else:
                try:
                    out = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,clocks.max.sm", "--format=csv,noheader,nounits", "-i",
                                          str(self.index)], capture_output=True, text=True, timeout=10).stdout.strip().split(",")
                    self.sm.append(float(out[0])); self.max_mhz = float(out[1])
                except Exception:  # noqa: BLE001
                    pass
                self.running.set()
                time.sleep(0.05)

    def summary(self):
        self.stop_flag = True
        names = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}
        reasons = [n for bit, n in names.items() if self.reason_bits & bit]
        sm = sorted(self.sm)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": self.max_mhz, "reasons": reasons, "samples": len(sm), "how": self.how}


def look_at(theta, phi, radius=4.0):
    eye = torch.tensor([radius * math.cos(phi) * math.cos(theta), radius * math.cos(phi) * math.sin(theta), radius * math.sin(phi)], dtype=torch.float64)
    back = eye / eye.norm()
    right = torch.linalg.cross(torch.tensor([0.0, 0.0, 1.0], dtype=torch.float64), back)
    right = right / right.norm()
    up = torch.linalg.cross(back, right)
    m = torch.eye(4, dtype=torch.float64)
    m[:3, 0], m[:3, 1], m[:3, 2], m[:3, 3] = right, up, back, eye
    return m.float()


def make_inputs(n_sets, rays, S, seed, pin=False, n_pixels=100 * 100, with_jitter=True):
    """synthetic step inputs on the HOST: pixel ids, target colours, (optionally) an explicit stratified-jitter tensor"""
    g = torch.Generator().manual_seed(seed)
    pix = torch.randint(0, n_pixels, (n_sets, rays), generator=g)
    tgt = torch.rand(n_sets, rays, 3, generator=g)
    jit = torch.rand(n_sets, rays, S, generator=g) if with_jitter else None
    if pin:
        pix, tgt = pix.pin_memory(), tgt.pin_memory()
        jit = jit.pin_memory() if jit is not None else None
    return pix, tgt, jit


def step_input_bytes(rays, S, jitter):
    """pixel ids (8 B) + targets (12 B) per ray, + 4*S B/ray when the jitter is an explicit tensor (default: drawn in-kernel, like
    the reference draws it on the device, src/sampling.py:24)"""
    return rays * (20 + (4 * S if jitter == "tensor" else 0))


def n_input_sets(rays, S, jitter="kernel"):
    return max(2, int(math.ceil(140e6 / step_input_bytes(rays, S, jitter))))       # > L2 (126 MB) of rotating inputs


def resolve(args):
    """workload -> shapes (shared by both arms)"""
    w = args.workload
    if w == "c4":
        args.samples, args.hidden = 192, 256
    elif w == "c5":
        args.samples = 128
        if args.rays is None:
            args.rays = 1 << 17
    elif w == "c1":
        args.samples = 64
        if args.rays is None:
            args.rays = 2048
    if args.rays is None:
        args.rays = 4096
    if args.rounds <= 0:
        args.rounds = max(5, min(50, int(math.ceil(1000.0 / max(1, args.steps)))))
    return args


def workload_config(args, world):
    """the `config` object of the JSON line: a function of the command line only, so both arms print the same dict"""
    S, rays, w = args.samples, args.rays, args.workload
    if w == "train":
        ns = n_input_sets(rays, S, args.jitter)
        return {"workload": f"C3 train.py random-ray batch: {rays} rays x {S} samples per GPU, fwd+bwd+Adam, L=10 hidden=128 depth=4 skip=2",
                "rays_per_gpu": rays, "samples": S, "parallelism": f"ray-sharded dp{world}",
                "jitter": "drawn on the device (GPU arm: in-kernel Philox; CPU arm: torch.rand_like as src/sampling.py:24)" if args.jitter == "kernel"
                          else "explicit (rays, samples) tensor",
                "l2": f"GPU arm: {ns} rotating input sets = {ns * step_input_bytes(rays, S, args.jitter) / 1e6:.0f} MB > 126 MB L2"}
    if w == "c1":
        ns = n_input_sets(rays, S, args.jitter)
        return {"workload": f"C1 tiny_nerf_min.py: train step {rays} rays x {S} samples fwd+bwd+Adam + 100x100 render_image, hidden=128 depth=4 skip=2, "
                            f"L=10 (as coded) and L=6 (as BASELINE words it); value = the L=10 train step",
                "rays_per_gpu": rays, "samples": S, "parallelism": f"replicas x{world}",
                "l2": f"GPU arm: train {ns} rotating input sets > 126 MB L2; render: 256 MB flush write between frames"}
    if w == "c5":
        return {"workload": f"C5 ray-batch sweep point: {rays} rays x {S} samples, fused fwd+bwd (no optimiser), L=10 hidden=128",
                "rays_per_gpu": rays, "samples": S, "parallelism": f"replicas x{world}",
                "l2": f"GPU arm: {n_input_sets(rays, S, args.jitter)} rotating input sets of {step_input_bytes(rays, S, args.jitter) / 1e6:.1f} MB > 126 MB L2"}
    if w == "c4":
        return {"workload": "C4 800x800 frame, 192 samples/ray, L=10 hidden=256, fused forward on CTA pairs; rows sharded over the ranks",
                "rays": 800 * 800, "rays_per_gpu": 800 * 800 // world, "samples": S, "parallelism": f"row-sharded x{world}",
                "l2": "GPU arm: 256 MB flush write between timed iterations"}
    return {"workload": "C2 full 100x100 view render, 64 samples/ray, fused forward, L=10 hidden=128", "rays": 100 * 100, "samples": S,
            "parallelism": f"replicas x{world}", "l2": "GPU arm: 256 MB flush write between timed iterations"}


# ----------------------------------------------------------------------------------------------------
# CPU side: the reference's own modules (oracle/_ref/src, unmodified) or the oracle port
def load_reference():
    d = os.path.join(ROOT, "oracle", "_ref", "src")
    names = ["rays", "sampling", "encoding", "nerf", "volume", "utils"]
    if not all(os.path.exists(os.path.join(d, n + ".py")) for n in names):
        return None
    mods = {}
    for n in names:                      # loaded under private names: the product package has modules of the same names
        spec = importlib.util.spec_from_file_location("tnerf_reference_" + n, os.path.join(d, n + ".py"))
        m = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(m)
        mods[n] = m
    return types.SimpleNamespace(**mods)


class CpuArm:
    """one training step (src/train.py:106-128) / one frame (src/train.py:36-59) on the host cores"""

    def __init__(self, hidden, L, S, H=100, W=100, focal=FOCAL, n_poses=8):
        self.R = load_reference()
        self.kind = "reference" if self.R is not None else "port"
        self.cores = os.cpu_count() or 1
        torch.set_num_threads(self.cores)
        self.S, self.H, self.W, self.focal = S, H, W, focal
        self.poses = [look_at(2 * math.pi * i / n_poses, 0.5) for i in range(n_poses)]
        torch.manual_seed(0)
        if self.R is not None:
            R = self.R
            self.enc = R.encoding.PositionalEncoding(L, True)
            self.model = R.nerf.TinyNeRF(self.enc.out_dim, hidden, 4, 2)
            self.opt = torch.optim.Adam(self.model.parameters(), lr=5e-4)
            self.rays = None
        else:
            from oracle import oracle as O
            self.O = O
            self.p = O.init_params(6 * L + 3, hidden, 4, 2, seed=0)
            self.L = L
            self.m = {k: torch.zeros_like(v) for k, v in self.p.items()}
            self.v = {k: torch.zeros_like(x) for k, x in self.p.items()}
        self.steps = 0

    def describe(self):
        return ("unmodified reference modules (oracle/_ref/src: rays, sampling, encoding, nerf, volume; torch CPU fp32, autograd + torch.optim.Adam)"
                if self.kind == "reference" else "oracle/oracle.py (torch CPU fp32 port)") + f", {self.cores} threads"

    def _all_rays(self):
        if self.rays is None:            # src/train.py:94-101 precomputes the rays of every pose once, outside the loop
            gr = self.R.rays.get_rays if self.R is not None else None
            ro, rd = zip(*[(gr(self.H, self.W, self.focal, p) if gr else self.O.get_rays(self.H, self.W, self.focal, p)) for p in self.poses])
            self.rays = (torch.stack([r.contiguous() for r in ro]), torch.stack(rd))
        return self.rays

    def train_step(self, i, pix, tgt, jit=None):
        ro_all, rd_all = self._all_rays()
        v = i % len(self.poses)
        ro, rd = ro_all[v, pix], rd_all[v, pix]
        self.steps += 1
        if self.R is not None:
            R, S, n = self.R, self.S, pix.shape[0]
            self.model.train()
            z_vals, pts = R.sampling.stratified_samples(2.0, 6.0, S, ro, rd, randomized=True)
            xenc = self.enc(pts.reshape(-1, 3))
            rgb, sigma = self.model(xenc)
            comp, _, _, _ = R.volume.volume_render(rgb.reshape(n, S, 3), sigma.reshape(n, S, 1), z_vals, rd)
            loss = torch.mean((comp - tgt) ** 2)
            R.utils.mse2psnr(loss)
            self.opt.zero_grad(set_to_none=True)
            loss.backward()
            self.opt.step()
            return loss
        O = self.O
        if jit is None:
            jit = torch.rand(pix.shape[0], self.S)
        loss, g, _ = O.loss_and_grads(self.p, ro, rd, tgt, 2.0, 6.0, self.S, jit, num_freqs=self.L) if "num_freqs" in O.loss_and_grads.__code__.co_varnames \
            else O.loss_and_grads(self.p, ro, rd, tgt, 2.0, 6.0, self.S, jit)
        O.adam_step(self.p, g, self.m, self.v, self.steps)
        return loss

    def fwd_bwd(self, i, pix, tgt, jit=None):
        """forward + backward without the optimiser (BASELINE config 5)"""
        ro_all, rd_all = self._all_rays()
        v = i % len(self.poses)
        ro, rd = ro_all[v, pix], rd_all[v, pix]
        if self.R is not None:
            R, S, n = self.R, self.S, pix.shape[0]
            z_vals, pts = R.sampling.stratified_samples(2.0, 6.0, S, ro, rd, randomized=True)
            rgb, sigma = self.model(self.enc(pts.reshape(-1, 3)))
            comp, _, _, _ = R.volume.volume_render(rgb.reshape(n, S, 3), sigma.reshape(n, S, 1), z_vals, rd)
            loss = torch.mean((comp - tgt) ** 2)
            self.opt.zero_grad(set_to_none=True)
            loss.backward()
            return loss
        if jit is None:
            jit = torch.rand(pix.shape[0], self.S)
        return self.O.loss_and_grads(self.p, ro, rd, tgt, 2.0, 6.0, self.S, jit)[0]

    @torch.no_grad()
    def render(self, i, lo, n, chunk=8192):
        """rays [lo, lo+n) of a frame, chunked like render_one (src/train.py:36-59)"""
        pose = self.poses[i % len(self.poses)]
        if self.R is not None:
            R, S = self.R, self.S
            self.model.eval()
            ro, rd = R.rays.get_rays(self.H, self.W, self.focal, pose)
            ro, rd = ro[lo:lo + n], rd[lo:lo + n]
            out = []
            for a in range(0, n, chunk):
                o, d = ro[a:a + chunk], rd[a:a + chunk]
                z_vals, pts = R.sampling.stratified_samples(2.0, 6.0, S, o, d, randomized=False)
                rgb, sigma = self.model(self.enc(pts.reshape(-1, 3)))
                comp, _, _, _ = R.volume.volume_render(rgb.reshape(pts.shape[0], S, 3), sigma.reshape(pts.shape[0], S, 1), z_vals, d)
                out.append(comp)
            return torch.cat(out, 0).clamp(0.0, 1.0)
        O = self.O
        ro, rd = O.get_rays(self.H, self.W, self.focal, pose)
        return O.render_rays(self.p, ro[lo:lo + n], rd[lo:lo + n], 2.0, 6.0, self.S, None)[0]


def cpu_measure(args, budget_s, steps=None, warmup=1):
    """(value, unit, sample text, arm) of the workload's step on the host cores, on a bounded sample of the workload"""
    w = args.workload
    c4 = w == "c4"
    S = args.samples
    if w in ("train", "c1", "c5"):
        arm = CpuArm(128, 10, S)
        rays_full = args.rays
        pix, tgt, jit = make_inputs(2, min(rays_full, 8192), S, 99)
        fn = arm.fwd_bwd if w == "c5" else arm.train_step
        rays = min(rays_full, 8192)
        t0 = time.perf_counter(); fn(0, pix[0, :rays], tgt[0, :rays], jit[0, :rays]); t1 = time.perf_counter() - t0
        n_steps = steps if steps is not None else 6
        while rays > 256 and t1 * (rays / min(rays_full, 8192)) * (n_steps + warmup) > budget_s:
            rays //= 2
        for i in range(warmup):
            fn(i, pix[i % 2, :rays], tgt[i % 2, :rays], jit[i % 2, :rays])
        t0 = time.perf_counter()
        n = 0
        while n < n_steps and (steps is not None or time.perf_counter() - t0 < budget_s):
            fn(n, pix[n % 2, :rays], tgt[n % 2, :rays], jit[n % 2, :rays]); n += 1
        dt = time.perf_counter() - t0
        what = "fwd+bwd" if w == "c5" else "fwd+bwd+Adam"
        return rays * S * n / dt, "ray-samples/s", f"{n} steps ({what}) of {rays} of {rays_full} rays x {S} samples; {arm.describe()}", arm, dt / n
    RH, RW_, rfocal = (800, 800, 1111.11) if c4 else (100, 100, FOCAL)
    arm = CpuArm(256 if c4 else 128, 10, S, RH, RW_, rfocal)
    rays_full = RH * RW_
    rays = 4096 if c4 else rays_full
    lo = (rays_full - rays) // 2                      # a slice from the middle of the frame (rays are independent)
    t0 = time.perf_counter(); arm.render(0, lo, rays); t1 = time.perf_counter() - t0
    n_steps = steps if steps is not None else 6
    while rays > 512 and t1 * (n_steps + warmup) > budget_s:
        rays //= 2; t1 /= 2
    lo = (rays_full - rays) // 2
    for i in range(warmup):
        arm.render(i, lo, rays)
    t0 = time.perf_counter()
    n = 0
    while n < n_steps and (steps is not None or time.perf_counter() - t0 < budget_s):
        arm.render(n, lo, rays); n += 1
    dt = time.perf_counter() - t0
    return rays * n / dt, "rays/s", f"{n} frames of {rays} of {rays_full} rays x {S} samples; {arm.describe()}", arm, dt / n


METRICS = {"train": ("train ray-samples/sec (fwd+bwd+Adam)", "ray-samples/s"), "c1": ("train ray-samples/sec (fwd+bwd+Adam)", "ray-samples/s"),
           "c5": ("train ray-samples/sec (fwd+bwd)", "ray-samples/s"), "render": ("render rays/sec (fused forward)", "rays/s"),
           "c4": ("render rays/sec (fused forward)", "rays/s")}


def reference_arm(args, rank, world):
    """CPU implementation of the same step on the host cores (rank 0 only; the other ranks exit without work)."""
    if rank != 0:
        return
    value, unit, sample, arm, sec = cpu_measure(args, budget_s=150.0, steps=args.steps, warmup=min(args.warmup, 2))
    extra = {}
    if args.workload == "c1":                     # the render_image half and the L=6 variant
        a6 = CpuArm(128, 6, args.samples)
        pix, tgt, jit = make_inputs(1, args.rays, args.samples, 7)
        a6.train_step(0, pix[0], tgt[0], jit[0])
        t0 = time.perf_counter()
        for i in range(2):
            a6.train_step(i, pix[0], tgt[0], jit[0])
        t6 = (time.perf_counter() - t0) / 2
        t0 = time.perf_counter(); arm.render(0, 0, 100 * 100); r10 = time.perf_counter() - t0
        t0 = time.perf_counter(); a6.render(0, 0, 100 * 100); r6 = time.perf_counter() - t0
        extra["variants"] = {"L10": {"train_ray_samples_per_s": value, "render_rays_per_s": 1e4 / r10},
                             "L6": {"train_ray_samples_per_s": args.rays * args.samples / t6, "render_rays_per_s": 1e4 / r6}}
    metric, _ = METRICS[args.workload]
    line = {"impl": "reference", "metric": metric, "value": value, "unit": unit, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * sec, "higher_is_better": True, "scaling": "strong" if args.workload == "c4" else "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload_config(args, world),
            "cpu_baseline": {"value": value, "unit": unit, "cores": arm.cores, "kind": arm.kind, "sample": sample},
            "e2e": {"value": value, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    line.update(extra)
    emit(line)


_STDOUT_FD = None


def quiet_stdout():
    """stdout carries ONE JSON line: libraries that write to file descriptor 1 (NCCL prints its version banner there on some boxes)
    are pointed at stderr for the duration of the run; emit() writes the line to the real stdout."""
    global _STDOUT_FD
    if _STDOUT_FD is None:
        sys.stdout.flush()
        _STDOUT_FD = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    sys.stdout.flush()
    if _STDOUT_FD is not None:
        os.write(_STDOUT_FD, (json.dumps(line) + "\n").encode())
    else:
        emit(line)


def cpu_baseline(args):
    value, unit, sample, arm, _ = cpu_measure(args, budget_s=20.0)
    return {"value": value, "unit": unit, "cores": arm.cores, "kind": arm.kind, "sample": sample}


# ----------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--rounds", type=int, default=0, help="timed rounds of --steps steps (median reported); 0 = about 1000 steps in total, 5..50 rounds")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="train", choices=["train", "render", "c1", "c4", "c5"])
    ap.add_argument("--rays", type=int, default=None, help="rays per GPU per step (train 4096, c1 2048, c5 131072)")
    ap.add_argument("--samples", type=int, default=64)
    ap.add_argument("--hidden", type=int, default=128)
    ap.add_argument("--sweep", action="store_true", help="c5: add every batch size 2^14..2^22 to the line")
    ap.add_argument("--jitter", default="kernel", choices=["kernel", "tensor"], help="stratified jitter: drawn in the training kernel (default) or an explicit tensor")
    ap.add_argument("--precision", default=None, help="f16 (tcgen05) or f32 (exact FFMA path); default: engine default")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-ddp-check", action="store_true")
    args = resolve(ap.parse_args())
    quiet_stdout()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        reference_arm(args, rank, world)
        return
    args.warmup = max(args.warmup, 3)

    if PKG not in sys.path:
        sys.path.insert(0, PKG)
    import ctypes as C
    import torch.distributed as dist
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    import _engine as E
    import engine
    from encoding import PositionalEncoding
    from nerf import TinyNeRF

    pk = peaks()
    K, R, W = args.steps, args.rounds, args.warmup
    poses = torch.stack([look_at(2 * math.pi * i / 106 + 0.01 * i, 0.25 + 0.6 * ((i * 37) % 106) / 106) for i in range(106)]).to(dev)
    c4 = args.workload == "c4"
    S, rays = args.samples, args.rays

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def timed_rounds(step, first, align=2, before_round=None, k=None, r=None):
        """R rounds of EXACTLY K steps, each bracketed by barrier + synchronize; `align` untimed steps after the barrier bring the
        ranks back together (the exchange kernel's rendezvous) before the start event.  Returns per-round ms, max over ranks."""
        k, r = k or K, r or R
        ev, i = [], first
        for _ in range(r):
            barrier()
            if before_round:
                before_round()
            for _a in range(align):
                step(i); i += 1
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _s in range(k):
                step(i); i += 1
            b.record()
            ev.append((a, b))
        barrier()
        t = torch.tensor([a.elapsed_time(b) for a, b in ev], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.cpu(), i

    def med(t):
        return float(t.sort().values[len(t) // 2])

    detail, extra = {}, {}
    ddp = None
    sampler = None

    def train_bench(L, n_rays, with_e2e=True, adam=True):
        """fused training step at (L, n_rays, S): device-resident rounds, end-to-end rounds, the dominant kernel alone"""
        nonlocal sampler
        torch.manual_seed(0)
        enc = PositionalEncoding(L, True).to(dev)
        model = TinyNeRF(enc.out_dim, 128, 4, 2).to(dev)
        tr = engine.Trainer(model, enc, n_samples=S, precision=args.precision)
        kj = args.jitter == "kernel"
        n_sets = n_input_sets(n_rays, S, args.jitter)
        pix_h, tgt_h, jit_h = make_inputs(n_sets, n_rays, S, 1234 + rank, pin=with_e2e, with_jitter=not kj)
        pix_d, tgt_d = pix_h.to(dev), tgt_h.to(dev)
        jit_d = None if kj else jit_h.to(dev)
        rs_cache = {}

        def fwd_bwd(i):
            k = i % n_sets
            key = (i % 106, k)
            if key not in rs_cache:
                rs_cache[key] = engine.ray_source(c2w=poses[i % 106], H=100, W=100, focal=FOCAL, pixel_index=pix_d[k],
                                                  jitter_seed=tr.jitter_seed if kj else 0)
            rs_cache[key].jitter_step = i
            E.check(E.lib().tnerf_train_fwd_bwd(tr.h.h, C.byref(rs_cache[key]), E.ptr(tgt_d[k]), n_rays, 2.0, 6.0, S, None if kj else E.ptr(jit_d[k]), 1,
                                                tr.prec, 3.0 * n_rays, None, E.ptr(tr.loss_view), E.ptr(tr.gbuf), None, None, E.stream(dev)))

        def step(i):
            k = i % n_sets
            return tr.step_pixels(poses[i % 106], 100, 100, FOCAL, pix_d[k], tgt_d[k], None if kj else jit_d[k], global_rays=n_rays * world)
        run = step if adam else fwd_bwd
        for i in range(W):
            run(i)
        if sampler is None:
            sampler = ClockSampler(local); sampler.start()
        l0 = E.launch_count()
        t_dev, nxt = timed_rounds(run, W)
        launches = (E.launch_count() - l0) / (R * (K + 2)) * K
        out = {"tr": tr, "t_dev": t_dev, "launches": launches, "n_sets": n_sets}
        if with_e2e:
            # ---- end to end: host (pinned) inputs in, loss out, every step.  The inputs of step i+1 are copied on a side stream
            #      while step i computes (two staging sets), the way a data loader feeds a training loop; every byte of every
            #      step still crosses PCIe inside the timed region and the loss of every step is read back.
            loss_h = torch.zeros(1).pin_memory()
            loss_slot = [torch.zeros(1, device=dev) for _ in range(2)]
            pose_h = poses.cpu().pin_memory()
            stage = [(torch.empty(4, 4, device=dev), torch.empty_like(pix_d[0]), torch.empty_like(tgt_d[0]), None if kj else torch.empty_like(jit_d[0]))
                     for _ in range(2)]
            copy_stream = torch.cuda.Stream(device=dev)
            ready = [torch.cuda.Event() for _ in range(2)]
            consumed = [torch.cuda.Event() for _ in range(2)]
            copied = torch.cuda.Event()
            main_stream = torch.cuda.current_stream(dev)
            copied.record(main_stream)

            def upload(i):
                k, slot = i % n_sets, i % 2
                with torch.cuda.stream(copy_stream):
                    copy_stream.wait_event(consumed[slot])
                    sb_pose, sb_pix, sb_tgt, sb_jit = stage[slot]
                    sb_pose.copy_(pose_h[i % 106], non_blocking=True)
                    sb_pix.copy_(pix_h[k], non_blocking=True); sb_tgt.copy_(tgt_h[k], non_blocking=True)
                    if not kj:
                        sb_jit.copy_(jit_h[k], non_blocking=True)
                    ready[slot].record(copy_stream)

            def step_e2e(i):
                slot = i % 2
                upload(i + 1)                                   # the copy of the NEXT step's inputs runs under this step
                main_stream.wait_event(ready[slot])
                main_stream.wait_event(copied)                  # the previous step's loss has left the buffer this step overwrites
                sb_pose, sb_pix, sb_tgt, sb_jit = stage[slot]
                o = tr.step_pixels(sb_pose, 100, 100, FOCAL, sb_pix, sb_tgt, sb_jit, global_rays=n_rays * world)
                consumed[slot].record(main_stream)
                # the loss of EVERY step is read back -- on the copy stream, behind this step's event, so that the compute stream keeps
                # its kernels back to back (a copy between them would break the programmatic dependent launches)
                with torch.cuda.stream(copy_stream):
                    copy_stream.wait_event(consumed[slot])
                    loss_slot[slot].copy_(o, non_blocking=True)
                    copied.record(copy_stream)
                    loss_h.copy_(loss_slot[slot], non_blocking=True)
            for evn in consumed:
                evn.record(main_stream)
            upload(nxt)
            t_e2e, _ = timed_rounds(step_e2e, nxt)
            out["t_e2e"] = t_e2e
            out["h2d"], out["d2h"] = 64 + step_input_bytes(n_rays, S, args.jitter), 4
            if med(t_dev) > med(t_e2e):                         # the end-to-end region does strictly more: re-measure the device-only rounds once
                t2, _ = timed_rounds(run, nxt + R * (K + 2) + 8)
                if med(t2) < med(t_dev):
                    out["t_dev"], out["remeasured"] = t2, True
        # ---- dominant kernel alone (roofline).  Optimised steps (adam=True) run the training kernel with the gradient left in its
        #      sum vector (no scatter launch), so that ONE launch is what is timed; the fwd+bwd-only workloads time the call a user
        #      of tnerf_train_fwd_bwd makes (training kernel + gradient scatter into the flat vector).
        solo = adam and tr.prec == E.PREC_F16_TC and tr.h.get_option("bulk_reduce") == 1 and os.environ.get("TNERF_GATHER", "1") != "0"
        if solo:
            E.check(E.lib().tnerf_set_sum_buffer(tr.h.h, None))

        def kernel_call(i):
            if not solo:
                return fwd_bwd(i)
            k = i % n_sets
            key = (i % 106, k)
            if key not in rs_cache:
                rs_cache[key] = engine.ray_source(c2w=poses[i % 106], H=100, W=100, focal=FOCAL, pixel_index=pix_d[k],
                                                  jitter_seed=tr.jitter_seed if kj else 0)
            rs_cache[key].jitter_step = i
            E.check(E.lib().tnerf_train_fwd_bwd(tr.h.h, C.byref(rs_cache[key]), E.ptr(tgt_d[k]), n_rays, 2.0, 6.0, S, None if kj else E.ptr(jit_d[k]), 1,
                                                tr.prec, 3.0 * n_rays, None, E.ptr(tr.loss_view), None, None, None, E.stream(dev)))
        reps = 20 if n_rays <= 65536 else 5
        lk0 = E.launch_count()
        for i in range(3):
            kernel_call(i)
        torch.cuda.synchronize()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        for i in range(reps):
            kernel_call(3 + i)
        g1.record()
        torch.cuda.synchronize()
        out["kernels_per_call"] = (E.launch_count() - lk0) / (reps + 3)
        out["ms_kernel"] = g0.elapsed_time(g1) / reps
        if solo:
            E.check(E.lib().tnerf_clear_sum(tr.h.h, E.stream(dev)))
        tr.gbuf.zero_()
        if tr.comm == "p2p":
            for v in tr._gviews:
                v.zero_()
        return out

    def render_bench(L, hidden, RH, RW_, rfocal, n, first):
        nonlocal sampler
        torch.manual_seed(0)
        enc = PositionalEncoding(L, True).to(dev)
        model = TinyNeRF(enc.out_dim, hidden, 4, 2).to(dev)
        h = E.handle_for(model, dev)
        h.set_encoding(L, True)
        prec = engine._PREC[args.precision] if args.precision else engine.default_precision()
        if prec == 0:
            h.ensure_packed(force=True)
        else:
            h.bind()
        comp, depth, acc = (torch.empty(n, 3, device=dev), torch.empty(n, 1, device=dev), torch.empty(n, 1, device=dev))
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

        def frame(i, pose=None):
            rs = engine.ray_source(c2w=poses[i % 106] if pose is None else pose, H=RH, W=RW_, focal=rfocal, first_ray=first)
            E.check(E.lib().tnerf_render_fwd(h.h, C.byref(rs), n, 2.0, 6.0, S, None, 1, prec, E.ptr(comp), E.ptr(depth), E.ptr(acc), None, None,
                                             E.stream(dev)))
        for i in range(W):
            frame(i)
        if sampler is None:
            sampler = ClockSampler(local); sampler.start()
        barrier()
        l0 = E.launch_count()
        evs = []
        for i in range(K):
            flush.zero_()                                   # L2 flush between timed iterations (not timed)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); frame(i); b.record()
            evs.append((a, b))
        barrier()
        ms = sum(a.elapsed_time(b) for a, b in evs)
        launches = E.launch_count() - l0
        # end to end: pose from pinned host memory in, full frame out to host
        pose_h = poses.cpu().pin_memory(); sb_pose = torch.empty(4, 4, device=dev)
        img_h = torch.empty(n, 3).pin_memory()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        f0.record()
        for i in range(K):
            sb_pose.copy_(pose_h[i % 106], non_blocking=True)
            frame(i, sb_pose)
            img_h.copy_(comp, non_blocking=True)
        f1.record()
        barrier()
        t = torch.tensor([ms, f0.elapsed_time(f1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return {"ms": float(t[0]), "ms_e2e": float(t[1]), "launches": launches, "prec": prec, "h2d": 64, "d2h": n * 12}

    prec_text = lambda p: "fp16 operands / fp32 accumulate (tcgen05)" if p == 0 else "fp32 FFMA"   # noqa: E731
    metric, unit = METRICS[args.workload]
    if args.workload in ("train", "c1", "c5"):
        adam = args.workload != "c5"
        if world > 1 and adam and not args.no_ddp_check:
            import ddp_train
            ddp = ddp_train.exchange_self_check(dev)          # multi-rank parity, in the warm-up (tests/test_gpu_ddp.py needs >= 2 GPUs)
        b = train_bench(10, rays, with_e2e=adam, adam=adam)
        tr = b["tr"]
        ms_round = med(b["t_dev"])
        ms_round_e2e = med(b["t_e2e"]) if adam else None
        units_per_step = rays * S
        flop = flop_per_sample(10, 128, True)
        ms_kernel, launches, dominant = b["ms_kernel"], b["launches"], "tnerf_train_fwd_bwd"
        h2d, d2h = (b["h2d"], b["d2h"]) if adam else (0, 0)
        detail = {"precision": prec_text(tr.prec), "kernels_per_fwd_bwd_call": b["kernels_per_call"], "exchange": tr.comm,
                  "rounds": R, "round_ms": {"median": ms_round, "min": float(b["t_dev"].min()), "max": float(b["t_dev"].max())},
                  "remeasured": bool(b.get("remeasured", False))}
        if tr.tile_order is not None:                       # SM-speed-aware tile dealing (engine.Trainer.calibrate_tile_order)
            detail["tile_order"] = engine._TILE_ORDERS[local][1]
        if adam:
            detail["round_ms_e2e"] = {"median": ms_round_e2e, "min": float(b["t_e2e"].min()), "max": float(b["t_e2e"].max())}
        if args.workload == "c5":
            # device-only arm of a forward+backward sweep point: the e2e leg uploads pixel ids / targets / jitter of every launch
            kj = args.jitter == "kernel"
            pix_h, tgt_h, jit_h = make_inputs(1, rays, S, 4321, pin=True, with_jitter=not kj)
            sb = (torch.empty_like(pix_h[0], device=dev), torch.empty_like(tgt_h[0], device=dev), None if kj else torch.empty_like(jit_h[0], device=dev))
            loss_h = torch.zeros(1).pin_memory()

            def e2e_step(i):
                sb[0].copy_(pix_h[0], non_blocking=True); sb[1].copy_(tgt_h[0], non_blocking=True)
                if not kj:
                    sb[2].copy_(jit_h[0], non_blocking=True)
                rs = engine.ray_source(c2w=poses[i % 106], H=100, W=100, focal=FOCAL, pixel_index=sb[0], jitter_seed=tr.jitter_seed if kj else 0, jitter_step=i)
                E.check(E.lib().tnerf_train_fwd_bwd(tr.h.h, C.byref(rs), E.ptr(sb[1]), rays, 2.0, 6.0, S, E.ptr(sb[2]), 1, tr.prec, 3.0 * rays, None,
                                                    E.ptr(tr.loss_view), E.ptr(tr.gbuf), None, None, E.stream(dev)))
                loss_h.copy_(tr.loss_view, non_blocking=True)
            ke = max(1, min(K, 20))
            t_e, _ = timed_rounds(e2e_step, 0, align=1, k=ke, r=3)
            ms_round_e2e = med(t_e) * K / ke
            h2d, d2h = step_input_bytes(rays, S, args.jitter), 4
            tr.gbuf.zero_()
            if args.sweep:
                rows = []
                for lg in range(14, 23):
                    n = 1 << lg
                    a2 = argparse.Namespace(**vars(args)); a2.rays = n
                    bb = train_bench(10, n, with_e2e=False, adam=False)
                    rows.append({"rays": n, "ms": bb["ms_kernel"], "ray_samples_per_s": n * S / (bb["ms_kernel"] * 1e-3),
                                 "frac_of_peak": n * S * flop / (bb["ms_kernel"] * 1e-3) / 1e12 / pk["tf_burst"]})
                    del bb
                    torch.cuda.empty_cache()
                extra["sweep"] = rows
        if args.workload == "c1":
            v = {"L10": {"train_ray_samples_per_s": world * units_per_step * K / (ms_round * 1e-3), "train_ms_per_step": ms_round / K,
                         "train_kernel_frac_of_peak": units_per_step * flop / (ms_kernel * 1e-3) / 1e12 / pk["tf_burst"]}}
            b6 = train_bench(6, rays, with_e2e=False)
            f6 = flop_per_sample(6, 128, True)
            v["L6"] = {"train_ray_samples_per_s": world * units_per_step * K / (med(b6["t_dev"]) * 1e-3), "train_ms_per_step": med(b6["t_dev"]) / K,
                       "train_kernel_frac_of_peak": units_per_step * f6 / (b6["ms_kernel"] * 1e-3) / 1e12 / pk["tf_burst"]}
            for L in (10, 6):
                r = render_bench(L, 128, 100, 100, FOCAL, 100 * 100, 0)
                v[f"L{L}"].update({"render_rays_per_s": world * 1e4 * K / (r["ms"] * 1e-3), "render_us_per_frame": 1e3 * r["ms"] / K,
                                   "render_e2e_rays_per_s": world * 1e4 * K / (r["ms_e2e"] * 1e-3),
                                   "render_frac_of_peak": 1e4 * S * flop_per_sample(L, 128, False) / (r["ms"] / K * 1e-3) / 1e12 / pk["tf_burst"]})
            extra["variants"] = v
    else:
        RH, RW_, rfocal = (800, 800, 1111.11) if c4 else (100, 100, FOCAL)
        n = RH * RW_ // world if c4 else RH * RW_            # C4: contiguous row block of the frame per rank; C2: one replica per rank
        r = render_bench(10, 256 if c4 else 128, RH, RW_, rfocal, n, rank * n if c4 else 0)
        ms_round, ms_round_e2e, launches = r["ms"], r["ms_e2e"], r["launches"]
        h2d, d2h = r["h2d"], r["d2h"]
        ms_kernel = ms_round / K
        units_per_step = n
        flop = flop_per_sample(10, 256 if c4 else 128, False) * S
        detail = {"precision": prec_text(r["prec"])}
        dominant = "tnerf_render_fwd_wide" if c4 else "tnerf_render_fwd"
    clocks = sampler.summary() if sampler is not None else None

    value = world * units_per_step * K / (ms_round * 1e-3)
    value_e2e = world * units_per_step * K / (ms_round_e2e * 1e-3)
    achieved_tf = units_per_step * flop / (ms_kernel * 1e-3) / 1e12
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        traffic = json.load(open(tpath)).get(dominant)
    line = {"metric": metric, "value": value, "unit": unit, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_round / K, "higher_is_better": True, "scaling": "strong" if c4 else "weak", "vs_baseline": None,
            "dtype": "f16" if "fp16" in detail["precision"] else "f32", "data": "synthetic", "config": workload_config(args, world),
            "e2e": {"value": value_e2e, "unit": unit, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": int(round(launches)), "clocks": clocks,
            "roofline": {"bound": "tensor", "kernel": dominant, "achieved": achieved_tf, "peak": pk["tf_burst"], "unit": "TFLOP/s",
                         "frac": achieved_tf / pk["tf_burst"], "traffic": traffic, "peak_source": pk["source"] + " bf16 burst (MEASURED_PEAKS.json)",
                         "frac_of_sustained_peak": achieved_tf / pk["tf_sustained"], "peak_sustained": pk["tf_sustained"],
                         "kernel_ms": ms_kernel, "flop_per_unit": flop},
            "detail": detail}
    line.update(extra)
    if ddp is not None:
        line["ddp_check"] = "ok" if ddp["ok"] else "FAIL"
        line["ddp_check_detail"] = {k: v for k, v in ddp.items() if k != "ok"}
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(args)
    if rank == 0:
        emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
