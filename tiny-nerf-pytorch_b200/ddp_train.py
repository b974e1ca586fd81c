"""ddp_train.py -- ray-sharded data-parallel training (the reference itself is single
process).  Launch one process per GPU:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 ddp_train.py [--iters ...]

Every rank holds the full MLP (265 KB), draws its own disjoint pixel / jitter streams, runs the fused
fwd+bwd kernel on its n_rand rays with the loss normalised by the GLOBAL ray count, and the flat
gradient (+loss) is summed AND the Adam step applied by ONE kernel over NVLink peer memory
(tnerf_allreduce_adam_step; TNERF_COMM=nccl selects all_reduce + tnerf_adam_step).  The sum is formed in
rank order, so every rank computes bit-identical parameters and they never need a broadcast.
"""
import os
from dataclasses import dataclass

import torch
import torch.distributed as dist

import engine
from data import load_tiny_nerf_npz
from encoding import PositionalEncoding
from nerf import TinyNeRF
from train import MODEL_CFG, render_one
from utils import mse2psnr


@dataclass
class Config:
    iters: int = 2000
    n_rand: int = 4096           # rays PER GPU per step (weak scaling)
    n_samples: int = 64
    lr: float = 5e-4
    near: float = 2.0
    far: float = 6.0
    log_every: int = 100
    data: str = "data/tiny_nerf_data.npz"
    ckpt_path: str = "checkpoints/tinynerf_latest.pth"


def shard_slices(n_total: int, world: int):
    """contiguous, balanced [begin, end) ray ranges, one per rank (used for sharded full-frame rendering)"""
    base, extra = divmod(n_total, world)
    out, a = [], 0
    for r in range(world):
        b = a + base + (1 if r < extra else 0)
        out.append((a, b))
        a = b
    return out


def rank_seed(base: int, rank: int, step: int = 0) -> int:
    """distinct, reproducible RNG streams per rank (pixel ids, jitter); identical model init everywhere"""
    return (base * 1000003 + rank * 7919 + step) % (2 ** 31 - 1)


def main(cfg: Config):
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    torch.manual_seed(0)                                   # same initial weights on every rank
    blob = load_tiny_nerf_npz(cfg.data)
    images = torch.from_numpy(blob["images"]).to(device)
    poses = torch.from_numpy(blob["poses"]).to(device)
    focal = float(blob["focal"])
    N, H, W, _ = images.shape
    pixels = images.view(N, H * W, 3)
    encoder = PositionalEncoding(10, True).to(device)
    model = TinyNeRF(encoder.out_dim, **MODEL_CFG).to(device)
    trainer = engine.Trainer(model, encoder, lr=cfg.lr, near=cfg.near, far=cfg.far, n_samples=cfg.n_samples)
    gen = torch.Generator(device=device).manual_seed(rank_seed(1234, rank))
    for step in range(cfg.iters):
        view = step % N
        pick = torch.randint(0, H * W, (cfg.n_rand,), device=device, generator=gen)
        # the stratified jitter is drawn inside the training kernel (per-rank Philox stream, engine.Trainer.jitter_seed)
        loss = trainer.step_pixels(poses[view], H, W, focal, pick, pixels[view].index_select(0, pick), None,
                                   global_rays=cfg.n_rand * world)
        if rank == 0 and (step + 1) % cfg.log_every == 0:
            lv = float(loss.item())
            print(f"[step {step + 1}] loss {lv:.5f} psnr {float(mse2psnr(torch.tensor(lv))):.2f} dB (global batch {cfg.n_rand * world} rays)")
    if rank == 0:
        os.makedirs(os.path.dirname(cfg.ckpt_path) or ".", exist_ok=True)
        torch.save({"model": model.state_dict(), "opt": trainer.state_dict(), "step": cfg.iters, "in_dim": encoder.out_dim,
                    "cfg": dict(MODEL_CFG)}, cfg.ckpt_path)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def exchange_self_check(device, steps: int = 4, n: int = 1024, n_samples: int = 64):
    """Multi-rank parity of the gradient exchange, callable from any initialised process group (bench.py runs it in the warm-up of
    every N > 1 line; tools/ddp_check.py and tests/test_gpu_ddp.py wrap it):
      * N-rank training on N ray shards == 1-rank training on the concatenated batch (up to fp32 summation order): every rank
        also trains alone on the whole batch and compares parameters and losses;
      * the parameters are BIT-IDENTICAL on every rank after the sharded steps (rank-ordered sum of the exchange kernel);
      * the peer-memory exchange and the NCCL exchange agree.
    Returns a dict with ``ok`` and the measured differences."""
    from encoding import PositionalEncoding
    from nerf import TinyNeRF
    rank, world = dist.get_rank(), dist.get_world_size()
    pose = torch.eye(4, device=device); pose[2, 3] = 4.0
    g = torch.Generator().manual_seed(77)
    pix_all = torch.randint(0, 10000, (steps, world * n), generator=g)
    tgt_all = torch.rand(steps, world * n, 3, generator=g)
    jit_all = torch.rand(steps, world * n, n_samples, generator=g)

    def run(comm, shard):
        torch.manual_seed(0)
        enc = PositionalEncoding(10, True).to(device)
        model = TinyNeRF(63, 128, 4, 2).to(device)
        if shard:
            tr = engine.Trainer(model, enc, n_samples=n_samples, comm=comm)
            sl = slice(rank * n, (rank + 1) * n)
        else:                                   # every rank trains alone on the whole batch (no exchange)
            tr = engine.Trainer(model, enc, n_samples=n_samples, comm="nccl")
            tr.world, tr.comm = 1, "none"
            sl = slice(0, world * n)
        losses = []
        for k in range(steps):
            out = tr.step_pixels(pose, 100, 100, 138.9, pix_all[k, sl].to(device), tgt_all[k, sl].to(device), jit_all[k, sl].to(device),
                                 global_rays=world * n)
            losses.append(out.clone())
        torch.cuda.synchronize(device)
        return tr.flat.clone(), torch.cat(losses), tr.comm

    p_p2p, l_p2p, used = run(None, True)         # the default exchange (peer memory when available)
    p_nccl, _, _ = run("nccl", True)
    p_one, l_one, _ = run("nccl", False)
    gathered = [torch.empty_like(p_p2p) for _ in range(world)]
    dist.all_gather(gathered, p_p2p)
    bit_identical = all(torch.equal(gathered[0], x) for x in gathered)
    def rel(a, b):
        return float((a - b).norm() / b.norm().clamp_min(1e-30))
    res = {"world": world, "comm": used, "steps": steps, "rays_per_rank": n, "bit_identical_across_ranks": bool(bit_identical),
           "rel_l2_p2p_vs_nccl": rel(p_p2p, p_nccl), "rel_l2_sharded_vs_single_rank": rel(p_p2p, p_one),
           "max_abs_sharded_vs_single_rank": float((p_p2p - p_one).abs().max()),
           "max_abs_loss_diff": float((l_p2p - l_one).abs().max()), "loss": float(l_p2p[-1])}
    # Adam divides by sqrt(v): an entry whose gradient is summation noise moves by up to lr per step in either direction, so the
    # LARGEST parameter difference between two summation orders is only bounded by steps * lr; the norm is the meaningful statement
    lr = 5e-4
    ok = (bit_identical and res["rel_l2_p2p_vs_nccl"] < 1e-4 and res["rel_l2_sharded_vs_single_rank"] < 1e-4
          and res["max_abs_sharded_vs_single_rank"] <= 2.1 * lr * steps and res["max_abs_loss_diff"] < 1e-3 * max(1.0, abs(res["loss"]))
          and bool(torch.isfinite(p_p2p).all()))
    flag = torch.tensor([1.0 if ok else 0.0], device=device)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)          # one verdict for the job
    res["ok"] = bool(flag.item() > 0.5)
    return res


@torch.no_grad()
def render_sharded(model, encoder, H, W, focal, pose, device, n_samples=64, near=2.0, far=6.0):
    """full frame with contiguous ray ranges per rank and one all_gather of the (rays/world, 3) shards"""
    import ctypes as C
    import _engine as E
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    a, b = shard_slices(H * W, world)[rank]
    h = E.handle_for(model, device)
    h.set_encoding(encoder.num_freqs, encoder.include_input)
    prec, _ = engine.pick_precisions(model, encoder, n_samples, device)
    if prec == E.PREC_F16_TC:
        h.ensure_packed()
    else:
        h.bind()
    pose_d = E.f32c(pose.to(device))
    part = torch.empty((b - a, 3), dtype=torch.float32, device=device)
    rs = engine.ray_source(c2w=pose_d, H=H, W=W, focal=focal, first_ray=a)
    E.check(E.lib().tnerf_render_fwd(h.h, C.byref(rs), b - a, near, far, n_samples, None, 1, prec, E.ptr(part), None, None, None, None,
                                     E.stream(device)), "tnerf_render_fwd")
    if world == 1:
        return part.reshape(H, W, 3).clamp(0, 1)
    sizes = [e - s for s, e in shard_slices(H * W, world)]
    pad = max(sizes)
    buf = torch.zeros((pad, 3), dtype=torch.float32, device=device)
    buf[: b - a] = part
    parts = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(parts, buf)
    return torch.cat([p[:n] for p, n in zip(parts, sizes)]).reshape(H, W, 3).clamp(0, 1)


if __name__ == "__main__":
    import tyro
    main(tyro.cli(Config))
