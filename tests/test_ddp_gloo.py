"""world_size-2 CPU (gloo) test of the ray-sharded data-parallel recipe used by ddp_train.py / engine.Trainer:
each rank differentiates its own ray shard with the loss normalised by the GLOBAL ray count, the flat
[gradient | loss] vector is all-reduced (sum), and the result must equal the single-process gradient of the
concatenated batch.  Kernels cannot run here, so the per-rank maths is the CPU oracle; what is under test is
the host-side contract (shard bookkeeping, normalisation, flat layout, one collective)."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tiny-nerf-pytorch_b200"))
    from oracle import oracle as O
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    p = O.init_params(27, 32, 4, 2, seed=0)
    g = torch.Generator().manual_seed(5)
    n_global, S = 64, 16
    pose = O.look_at_pose(0.4, 0.5)
    ro, rd = O.get_rays(16, 16, 20.0, pose)
    pix = torch.randint(0, 256, (n_global,), generator=g)
    u, tgt = torch.rand(n_global, S, generator=g), torch.rand(n_global, 3, generator=g)
    names = [k for k, _ in O.mlp_param_shapes(27, 32, 4, 2)]
    per = n_global // world
    sl = slice(rank * per, (rank + 1) * per)
    loss, grads, _ = O.loss_and_grads(p, ro[pix[sl]], rd[pix[sl]], tgt[sl], 2.0, 6.0, S, u[sl], num_freqs=4, denom=3 * n_global)
    flat = torch.cat([grads[k].reshape(-1) for k in names] + [loss.reshape(1)])
    dist.all_reduce(flat)                                        # the ONE collective of a training step
    m = {k: torch.zeros_like(v) for k, v in p.items()}
    v = {k: torch.zeros_like(x) for k, x in p.items()}
    off = 0
    red = {}
    for k in names:
        red[k] = flat[off:off + p[k].numel()].view_as(p[k]); off += p[k].numel()
    O.adam_step(p, red, m, v, 1)
    if rank == 0:
        ref_loss, ref_grads, _ = O.loss_and_grads(O.init_params(27, 32, 4, 2, seed=0), ro[pix], rd[pix], tgt, 2.0, 6.0, S, u, num_freqs=4)
        ref_flat = torch.cat([ref_grads[k].reshape(-1) for k in names] + [ref_loss.reshape(1)])
        out["grad_err"] = ((flat - ref_flat).norm() / ref_flat.norm()).item()
    # every rank applied the same update without a broadcast
    chk = torch.cat([p[k].reshape(-1) for k in names])
    lo, hi = chk.clone(), chk.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    if rank == 0:
        out["param_spread"] = (hi - lo).abs().max().item()
    dist.destroy_process_group()


def test_sharded_gradient_equals_global_batch_gradient():
    import socket
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    assert out["grad_err"] < 1e-5, dict(out)
    assert out["param_spread"] == 0.0, dict(out)


def test_shard_slices_and_rank_seeds():
    sys.path.insert(0, os.path.join(ROOT, "tiny-nerf-pytorch_b200"))
    import ddp_train
    for n, w in ((640000, 8), (10000, 3), (7, 8), (0, 2)):
        sl = ddp_train.shard_slices(n, w)
        assert len(sl) == w and sl[0][0] == 0 and sl[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(sl, sl[1:]))
        sizes = [b - a for a, b in sl]
        assert max(sizes) - min(sizes) <= 1
    seeds = {ddp_train.rank_seed(1234, r) for r in range(8)}
    assert len(seeds) == 8
