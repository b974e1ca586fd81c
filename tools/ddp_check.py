"""Two-or-more-rank check of the gradient exchange (run under torchrun on a multi-GPU box): ddp_train.exchange_self_check --
N-rank training on N shards == 1-rank training on the concatenated batch, parameters bit-identical across ranks, peer-memory
exchange == NCCL exchange.  Prints one line 'DDP_CHECK OK ...' on rank 0.  (bench.py runs the same check in every N > 1 line.)"""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tiny-nerf-pytorch_b200"))
import ddp_train  # noqa: E402

rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
res = ddp_train.exchange_self_check(dev, steps=4)
ok = res["ok"] and (res["comm"] == "p2p" or os.environ.get("TNERF_COMM") == "nccl")
if rank == 0:
    print(f"DDP_CHECK {'OK' if ok else 'FAIL'} " + " ".join(f"{k}={v}" for k, v in res.items()), flush=True)
dist.destroy_process_group()
sys.exit(0 if ok else 1)
