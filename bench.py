#!/usr/bin/env python
"""bench.py -- headline measurement of the TinyNeRF ray engine (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload train|render] [--impl reference]

Default workload = BASELINE config 3: train.py's random-ray batch, 4096 rays x 64 samples per GPU,
fused fwd+bwd + Adam (L=10, hidden 128, depth 4, skip 2), ray-sharded data parallel (weak scaling)
with one all-reduce of the 265 KB gradient.  `--workload render` times BASELINE config 2 (full
100x100 view, 64 samples, fused forward) and reports rays/s.  One JSON line is printed by rank 0.

`--impl reference` times the reference algorithm's CPU implementation (oracle/oracle.py: the same
ATen CPU ops the reference's PyTorch path executes; the reference itself is Python and cannot travel
to the GPU box) on the host cores with the same config / metric.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "tiny-nerf-pytorch_b200")
for _p in (ROOT, PKG):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import torch  # noqa: E402

FLOP_FWD = 131584          # 2*MAC of the Linear layers, L=10, hidden 128
FLOP_FWD_256 = 459776      # same, hidden 256
FLOP_FWD_BWD = 362496
FOCAL = 138.88888549804688


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return {"hbm_gbs": d["hbm_gbs"], "tf_burst": d["bf16_tflops"], "tf_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                "source": "measured"}
    return {"hbm_gbs": 6650.0, "tf_burst": 1590.0, "tf_sustained": 1400.0, "source": "fallback"}


class ClockSampler(threading.Thread):
    """SM clock and throttle reasons sampled DURING the timed region (B200_PROFILING.md), in-process through NVML
    (nvidia-ml-py) every ~2 ms; falls back to polling nvidia-smi."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.sm, self.reason_bits, self.stop_flag, self.max_mhz, self.how = index, [], 0, False, None, "nvml"
        self.nv = None
        try:
            import pynvml
            pynvml.nvmlInit()
            uuid = None
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
        """start the sampling thread and return only once it is running (short timed regions would otherwise end before
        the first sample)"""
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


def make_inputs(n_sets, rays, S, seed, pin=False):
    """synthetic step inputs on the HOST: pose id, pixel ids, target colours, stratified jitter"""
    g = torch.Generator().manual_seed(seed)
    pix = torch.randint(0, 100 * 100, (n_sets, rays), generator=g)
    tgt = torch.rand(n_sets, rays, 3, generator=g)
    jit = torch.rand(n_sets, rays, S, generator=g)
    if pin:
        pix, tgt, jit = pix.pin_memory(), tgt.pin_memory(), jit.pin_memory()
    return pix, tgt, jit


# ----------------------------------------------------------------------------------------------------
def reference_arm(args, rank, world):
    """CPU implementation of the same step on the host cores (rank 0 only)."""
    if rank != 0:
        return
    from oracle import oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    c4 = args.workload == "c4"
    S = 192 if c4 else args.samples
    p = O.init_params(63, 256 if c4 else 128, 4, 2, seed=0)
    poses = [look_at(2 * math.pi * i / 8, 0.5) for i in range(8)]
    if args.workload == "train":
        rays_full = args.rays
        pix, tgt, jit = make_inputs(4, rays_full, S, 1234)
        m = {k: torch.zeros_like(v) for k, v in p.items()}
        v = {k: torch.zeros_like(x) for k, x in p.items()}

        def one(i, rays):
            ro, rd = O.get_rays(100, 100, FOCAL, poses[i % 8])
            idx = pix[i % 4, :rays]
            _, g, _ = O.loss_and_grads(p, ro[idx], rd[idx], tgt[i % 4, :rays], 2.0, 6.0, S, jit[i % 4, :rays])
            O.adam_step(p, g, m, v, i + 1)
        t0 = time.perf_counter(); one(0, rays_full); t1 = time.perf_counter() - t0
        total = args.steps + args.warmup
        rays = rays_full
        while rays > 256 and t1 * (rays / rays_full) * total > 150.0:
            rays //= 2
        for i in range(args.warmup):
            one(i, rays)
        t0 = time.perf_counter()
        for i in range(args.steps):
            one(i, rays)
        dt = time.perf_counter() - t0
        value = rays * S * args.steps / dt
        unit, metric = "ray-samples/s", "train ray-samples/sec (fwd+bwd+Adam)"
        sample = f"{rays} of {rays_full} rays x {S} samples per step, {args.steps} steps, fp32, torch CPU ({cores} threads)"
        cfg = {"workload": "C3 train.py random-ray batch 4096x64 fwd+bwd+Adam", "rays_per_gpu": rays_full, "samples": S}
    else:
        RH, RW_, rfocal = (800, 800, 1111.11) if c4 else (100, 100, FOCAL)
        rays_full = RH * RW_

        def one(i, rays):
            ro, rd = O.get_rays(RH, RW_, rfocal, poses[i % 8])
            lo = (rays_full - rays) // 2              # a slice from the middle of the frame (rays are independent)
            with torch.no_grad():
                O.render_rays(p, ro[lo:lo + rays], rd[lo:lo + rays], 2.0, 6.0, S, None)
        rays = 4096 if c4 else rays_full
        t0 = time.perf_counter(); one(0, rays); t1 = time.perf_counter() - t0
        while rays > 512 and t1 * (args.steps + args.warmup) > 150.0:
            rays //= 2; t1 /= 2
        for i in range(args.warmup):
            one(i, rays)
        t0 = time.perf_counter()
        for i in range(args.steps):
            one(i, rays)
        dt = time.perf_counter() - t0
        value = rays * args.steps / dt
        unit, metric = "rays/s", "render rays/sec (fused forward)"
        sample = f"{rays} of {rays_full} rays x {S} samples per step, {args.steps} steps, fp32, torch CPU ({cores} threads)"
        cfg = {"workload": "C2 full 100x100 view render, 64 samples/ray", "rays": rays_full, "samples": S}
        if c4:
            cfg["workload"] = "C4 800x800 frame, 192 samples/ray, L=10 hidden=256"
    line = {"impl": "reference", "metric": metric, "value": value, "unit": unit, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "strong" if args.workload == "c4" else "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": cfg,
            "cpu_baseline": {"value": value, "unit": unit, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------------
def cpu_baseline(args):
    from oracle import oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    c4 = args.workload == "c4"
    S = 192 if c4 else args.samples
    p = O.init_params(63, 256 if c4 else 128, 4, 2, seed=0)
    pose = look_at(0.3, 0.5)
    ro, rd = O.get_rays(800, 800, 1111.11, pose) if c4 else O.get_rays(100, 100, FOCAL, pose)
    if c4:
        ro, rd = ro[318000:320048], rd[318000:320048]          # 2048 rays from the middle of the frame
    if args.workload == "train":
        rays = min(args.rays, 4096)
        pix, tgt, jit = make_inputs(1, rays, S, 99)
        m = {k: torch.zeros_like(v) for k, v in p.items()}
        v = {k: torch.zeros_like(x) for k, x in p.items()}

        def one(i):
            _, g, _ = O.loss_and_grads(p, ro[pix[0]], rd[pix[0]], tgt[0], 2.0, 6.0, S, jit[0])
            O.adam_step(p, g, m, v, i + 1)
        units, unit, reps = rays * S, "ray-samples/s", 6
    else:
        rays = int(ro.shape[0])

        def one(i):
            with torch.no_grad():
                O.render_rays(p, ro, rd, 2.0, 6.0, S, None)
        units, unit, reps = rays, "rays/s", 6
    one(0)
    t0 = time.perf_counter()
    n = 0
    while n < reps and time.perf_counter() - t0 < 25.0:
        one(n + 1); n += 1
    dt = time.perf_counter() - t0
    return {"value": units * n / dt, "unit": unit, "cores": cores, "kind": "port",
            "sample": f"{n} steps of {rays} rays x {S} samples, oracle/oracle.py (torch CPU fp32, {cores} threads)"}


# ----------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="train", choices=["train", "render", "c4"])
    ap.add_argument("--rays", type=int, default=4096, help="rays per GPU per train step")
    ap.add_argument("--samples", type=int, default=64)
    ap.add_argument("--precision", default=None, help="f16 (tcgen05) or f32 (exact FFMA path); default: engine default")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        reference_arm(args, rank, world)
        return

    import torch.distributed as dist
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    import _engine as E
    import engine
    from encoding import PositionalEncoding
    from nerf import TinyNeRF

    torch.manual_seed(0)
    enc = PositionalEncoding(10, True).to(dev)
    c4 = args.workload == "c4"           # BASELINE config 4: 800x800 frame, 192 samples/ray, hidden 256, rows sharded over the ranks
    model = TinyNeRF(enc.out_dim, 256 if c4 else 128, 4, 2).to(dev)
    if c4:
        args.samples = 192
    S, rays = args.samples, args.rays
    poses = torch.stack([look_at(2 * math.pi * i / 106 + 0.01 * i, 0.25 + 0.6 * ((i * 37) % 106) / 106) for i in range(106)]).to(dev)
    pk = peaks()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    if args.workload == "train":
        tr = engine.Trainer(model, enc, n_samples=S, precision=args.precision)
        n_sets = max(8, int(math.ceil(140e6 / (rays * (S * 4 + 20)))))       # > L2 (126 MB) of rotating inputs
        pix_h, tgt_h, jit_h = make_inputs(n_sets, rays, S, 1234 + rank, pin=True)
        pix_d, tgt_d, jit_d = pix_h.to(dev), tgt_h.to(dev), jit_h.to(dev)

        def step(i):
            k = i % n_sets
            return tr.step_pixels(poses[i % 106], 100, 100, FOCAL, pix_d[k], tgt_d[k], jit_d[k], global_rays=rays * world)
        for i in range(args.warmup):
            step(i)
        barrier()
        sampler = ClockSampler(local); sampler.start()
        l0 = E.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(args.steps):
            loss = step(args.warmup + i)
        e1.record()
        barrier()
        launches = E.launch_count() - l0
        clocks = sampler.summary()
        ms = e0.elapsed_time(e1)
        # ---- end-to-end: host (pinned) inputs in, loss out, every step.  The inputs of step i+1 are copied on a side stream
        #      while step i computes (two staging sets), the way a data loader feeds a training loop; every byte of every
        #      step still crosses PCIe inside the timed region and the loss of every step is read back.
        loss_h = torch.zeros(1).pin_memory()
        pose_h = poses.cpu().pin_memory()
        stage = [(torch.empty(4, 4, device=dev), torch.empty_like(pix_d[0]), torch.empty_like(tgt_d[0]), torch.empty_like(jit_d[0])) for _ in range(2)]
        copy_stream = torch.cuda.Stream(device=dev)
        ready = [torch.cuda.Event() for _ in range(2)]
        consumed = [torch.cuda.Event() for _ in range(2)]
        main_stream = torch.cuda.current_stream(dev)

        def upload(i):
            k, slot = i % n_sets, i % 2
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(consumed[slot])
                sb_pose, sb_pix, sb_tgt, sb_jit = stage[slot]
                sb_pose.copy_(pose_h[i % 106], non_blocking=True)
                sb_pix.copy_(pix_h[k], non_blocking=True); sb_tgt.copy_(tgt_h[k], non_blocking=True); sb_jit.copy_(jit_h[k], non_blocking=True)
                ready[slot].record(copy_stream)

        def step_e2e(i, last):
            slot = i % 2
            if not last:
                upload(i + 1)
            main_stream.wait_event(ready[slot])
            sb_pose, sb_pix, sb_tgt, sb_jit = stage[slot]
            out = tr.step_pixels(sb_pose, 100, 100, FOCAL, sb_pix, sb_tgt, sb_jit, global_rays=rays * world)
            consumed[slot].record(main_stream)
            loss_h.copy_(out, non_blocking=True)
        for ev in consumed:
            ev.record(main_stream)
        upload(0)
        for i in range(3):
            step_e2e(i, i == 2)
        barrier()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record()
        upload(3)                                           # all K uploads of the K timed steps are issued inside the timed region
        for i in range(args.steps):
            step_e2e(3 + i, i == args.steps - 1)
        f1.record()
        barrier()
        ms_e2e = f0.elapsed_time(f1)
        h2d = 64 + rays * 8 + rays * 12 + rays * S * 4
        d2h = 4
        # ---- dominant kernel alone (roofline): the fused fwd+bwd launch, no optimiser
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 20
        rs_list = [engine.ray_source(c2w=poses[i % 106], H=100, W=100, focal=FOCAL, pixel_index=pix_d[i % n_sets]) for i in range(reps)]
        import ctypes as C
        lk0 = E.launch_count()
        torch.cuda.synchronize()
        g0.record()
        for i in range(reps):
            k = i % n_sets
            E.check(E.lib().tnerf_train_fwd_bwd(tr.h.h, C.byref(rs_list[i]), E.ptr(tgt_d[k]), rays, 2.0, 6.0, S, E.ptr(jit_d[k]), 1, tr.prec,
                                                3.0 * rays, None, E.ptr(tr.loss_view), E.ptr(tr.gbuf), E.stream(dev)))
        g1.record()
        torch.cuda.synchronize()
        kernels_per_call = (E.launch_count() - lk0) / reps
        tr.gbuf.zero_()
        ms_kernel = g0.elapsed_time(g1) / reps
        units_per_step = rays * S
        flop = FLOP_FWD_BWD
        unit, metric = "ray-samples/s", "train ray-samples/sec (fwd+bwd+Adam)"
        cfg = {"workload": "C3 train.py random-ray batch: 4096 rays x 64 samples per GPU, fwd+bwd+Adam, L=10 hidden=128 depth=4 skip=2",
               "rays_per_gpu": rays, "samples": S, "parallelism": f"ray-sharded dp{world}",
               "precision": "fp16 operands / fp32 accumulate (tcgen05)" if tr.prec == 0 else "fp32 FFMA",
               "l2": f"{n_sets} rotating input sets = {n_sets * rays * (S * 4 + 20) / 1e6:.0f} MB > 126 MB L2", "kernels_per_fwd_bwd_call": kernels_per_call}
        dominant = "tnerf_train_fwd_bwd"
    else:
        h = E.handle_for(model, dev)
        h.set_encoding(10, True)
        prec = engine._PREC[args.precision] if args.precision else engine.default_precision()
        if prec == 0:
            h.ensure_packed(force=True)
        RH, RW_, rfocal = (800, 800, 1111.11) if c4 else (100, 100, FOCAL)
        n = RH * RW_ // world if c4 else RH * RW_            # C4: contiguous row block of the frame per rank; C2: one replica per rank
        first = rank * n if c4 else 0
        comp, depth, acc = (torch.empty(n, 3, device=dev), torch.empty(n, 1, device=dev), torch.empty(n, 1, device=dev))
        import ctypes as C
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

        def step(i):
            rs = engine.ray_source(c2w=poses[i % 106], H=RH, W=RW_, focal=rfocal, first_ray=first)
            E.check(E.lib().tnerf_render_fwd(h.h, C.byref(rs), n, 2.0, 6.0, S, None, 1, prec, E.ptr(comp), E.ptr(depth), E.ptr(acc), None, None,
                                             E.stream(dev)))
        for i in range(args.warmup):
            step(i)
        barrier()
        sampler = ClockSampler(local); sampler.start()
        l0 = E.launch_count()
        ms = 0.0
        evs = []
        for i in range(args.steps):
            flush.zero_()                                   # L2 flush between timed iterations (not timed)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); step(i); b.record()
            evs.append((a, b))
        barrier()
        ms = sum(a.elapsed_time(b) for a, b in evs)
        launches = E.launch_count() - l0
        clocks = sampler.summary()
        # end to end: pose from pinned host memory in, full frame out to host
        pose_h = poses.cpu().pin_memory(); sb_pose = torch.empty(4, 4, device=dev)
        img_h = torch.empty(n, 3).pin_memory()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        f0.record()
        for i in range(args.steps):
            sb_pose.copy_(pose_h[i % 106], non_blocking=True)
            rs = engine.ray_source(c2w=sb_pose, H=RH, W=RW_, focal=rfocal, first_ray=first)
            E.check(E.lib().tnerf_render_fwd(h.h, C.byref(rs), n, 2.0, 6.0, S, None, 1, prec, E.ptr(comp), E.ptr(depth), E.ptr(acc), None, None,
                                             E.stream(dev)))
            img_h.copy_(comp, non_blocking=True)
        f1.record()
        barrier()
        ms_e2e = f0.elapsed_time(f1)
        h2d, d2h = 64, n * 12
        ms_kernel = ms / args.steps
        units_per_step = n
        flop = (FLOP_FWD_256 if c4 else FLOP_FWD) * S
        unit, metric = "rays/s", "render rays/sec (fused forward)"
        cfg = {"workload": "C2 full 100x100 view render, 64 samples/ray, fused forward, L=10 hidden=128", "rays": n, "samples": S,
               "parallelism": f"replicas x{world}", "precision": "fp16 operands / fp32 accumulate (tcgen05)" if prec == 0 else "fp32 FFMA",
               "l2": "256 MB flush write between timed iterations"}
        if c4:
            cfg.update({"workload": "C4 800x800 frame, 192 samples/ray, L=10 hidden=256, fused forward on CTA pairs; rows sharded over the ranks",
                        "rays": RH * RW_, "rays_per_gpu": n, "parallelism": f"row-sharded x{world}"})
        dominant = "tnerf_render_fwd_wide" if c4 else "tnerf_render_fwd"

    t = torch.tensor([ms, ms_e2e], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ms_e2e = t.tolist()
    value = world * units_per_step * args.steps / (ms * 1e-3)
    value_e2e = world * units_per_step * args.steps / (ms_e2e * 1e-3)
    achieved_tf = units_per_step * flop / (ms_kernel * 1e-3) / 1e12
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        traffic = json.load(open(tpath)).get(dominant)
    line = {"metric": metric, "value": value, "unit": unit, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong" if c4 else "weak", "vs_baseline": None,
            "dtype": "f16" if "fp16" in cfg["precision"] else "f32", "data": "synthetic", "config": cfg,
            "e2e": {"value": value_e2e, "unit": unit, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": int(launches), "clocks": clocks,
            "roofline": {"bound": "tensor", "kernel": dominant, "achieved": achieved_tf, "peak": pk["tf_burst"], "unit": "TFLOP/s",
                         "frac": achieved_tf / pk["tf_burst"], "traffic": traffic, "peak_source": pk["source"] + " bf16 burst (MEASURED_PEAKS.json)",
                         "kernel_ms": ms_kernel, "flop_per_unit": flop}}
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(args)
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
