"""Two-or-more-rank check of the gradient exchange (run under torchrun on a multi-GPU box):
  * "p2p" (tnerf_allreduce_adam_step: one kernel, NVLink peer memory) and "nccl" (all_reduce + tnerf_adam_step) must give the
    same parameters after K steps up to summation order, and the same loss;
  * with p2p the parameters must be BIT-IDENTICAL on every rank (rank-ordered sum);
  * N-rank training on N shards == 1-rank training on the concatenated batch (up to fp32 summation order).
Prints one line 'DDP_CHECK OK ...' on rank 0."""
import math
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tiny-nerf-pytorch_b200"))
import engine  # noqa: E402
from encoding import PositionalEncoding  # noqa: E402
from nerf import TinyNeRF  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
K, n, S = 6, 1024, 64
pose = torch.eye(4, device=dev); pose[2, 3] = 4.0
g = torch.Generator().manual_seed(77)
pix_all = torch.randint(0, 10000, (K, world * n), generator=g)
tgt_all = torch.rand(K, world * n, 3, generator=g)
jit_all = torch.rand(K, world * n, S, generator=g)


def run(comm, shard):
    torch.manual_seed(0)
    enc = PositionalEncoding(10, True).to(dev)
    model = TinyNeRF(63, 128, 4, 2).to(dev)
    if shard:
        tr = engine.Trainer(model, enc, n_samples=S, comm=comm)
        sl = slice(rank * n, (rank + 1) * n)
    else:                                   # every rank trains alone on the whole batch (no exchange)
        tr = engine.Trainer(model, enc, n_samples=S, comm="nccl")
        tr.world, tr.comm = 1, "none"
        sl = slice(0, world * n)
    losses = []
    for k in range(K):
        out = tr.step_pixels(pose, 100, 100, 138.9, pix_all[k, sl].to(dev), tgt_all[k, sl].to(dev), jit_all[k, sl].to(dev), global_rays=world * n)
        losses.append(out.clone())
    torch.cuda.synchronize()
    return tr.flat.clone(), torch.cat(losses), tr.comm


p_p2p, l_p2p, used = run("p2p", True)
p_nccl, l_nccl, _ = run("nccl", True)
p_one, l_one, _ = run("nccl", False)
gathered = [torch.empty_like(p_p2p) for _ in range(world)]
dist.all_gather(gathered, p_p2p)
bit_identical = all(torch.equal(gathered[0], x) for x in gathered)
d_modes = (p_p2p - p_nccl).abs().max().item()
d_single = (p_p2p - p_one).abs().max().item()
dl = (l_p2p - l_one).abs().max().item()
ok = used == "p2p" and bit_identical and d_modes < 5e-5 and d_single < 2e-4 and dl < 1e-5 and bool(torch.isfinite(p_p2p).all())
if rank == 0:
    print(f"DDP_CHECK {'OK' if ok else 'FAIL'} world={world} comm={used} bit_identical_across_ranks={bit_identical} "
          f"max|p2p-nccl|={d_modes:.2e} max|sharded-single|={d_single:.2e} max|loss diff|={dl:.2e} loss={l_p2p.tolist()[-1]:.5f}", flush=True)
dist.destroy_process_group()
sys.exit(0 if ok else 1)
