"""Time of the peer-memory all-reduce + Adam launch alone (no training kernel in between), per rank count; run under torchrun.
Compares against the single-GPU optimiser launch and the NCCL all-reduce of the same vector.  Developer tool (GPU box)."""
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
torch.manual_seed(0)
enc = PositionalEncoding(10, True).to(dev)
res = {}
for comm in ("p2p", "nccl"):
    model = TinyNeRF(63, 128, 4, 2).to(dev)
    tr = engine.Trainer(model, enc, n_samples=64, comm=comm)
    for reps in (20, 200):
        torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            tr._finish()
        b.record()
        torch.cuda.synchronize()
        t = torch.tensor([a.elapsed_time(b) / reps * 1e3], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    res[comm] = float(t)
tr1 = engine.Trainer(TinyNeRF(63, 128, 4, 2).to(dev), enc, n_samples=64)
tr1.world, tr1.comm = 1, "none"
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(200):
    tr1._finish()
b.record()
torch.cuda.synchronize()
if rank == 0:
    print(f"EXCHANGE world={world}: peer-memory all-reduce+Adam {res['p2p']:.1f} us/call, NCCL all-reduce + optimiser launch {res['nccl']:.1f} us/call, "
          f"single-GPU optimiser launch {a.elapsed_time(b) / 200 * 1e3:.1f} us/call", flush=True)
dist.destroy_process_group()
