"""Per-tensor gradient errors of the drop-in autograd path on the smoke() case (developer tool)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tiny-nerf-pytorch_b200"))
from oracle import oracle as O
import _engine as E, engine
from encoding import PositionalEncoding
from nerf import TinyNeRF
from rays import get_rays
dev = torch.device("cuda:0")
torch.manual_seed(0)
enc = PositionalEncoding(10, True).to(dev)
model = TinyNeRF(enc.out_dim, 128, 4, 2).to(dev)
pose = O.look_at_pose(0.6, 0.5, 4.0)
H = W = int(sys.argv[1]) if len(sys.argv) > 1 else 32
focal = 44.0
ro, rd = get_rays(H, W, focal, pose.to(dev))
n, S = H * W, 64
u = torch.rand(n, S, generator=torch.Generator().manual_seed(1))
target = torch.rand(n, 3, generator=torch.Generator().manual_seed(2))
for rep in range(3):
    model.zero_grad(set_to_none=True)
    comp, depth, acc = engine.render_rays(model, enc, ro, rd, 2.0, 6.0, S, t_rand=u.to(dev))
    loss = ((comp - target.to(dev)) ** 2).mean()
    loss.backward()
    torch.cuda.synchronize()
    p = {k: v.detach().cpu() for k, v in model.state_dict().items()}
    oro, ord_ = O.get_rays(H, W, focal, pose)
    l_ref, g_ref, _ = O.loss_and_grads(p, oro, ord_, target, 2.0, 6.0, S, u)
    print("rep", rep, "loss", loss.item(), l_ref.item())
    for k, g in g_ref.items():
        gg = dict(model.named_parameters())[k].grad.cpu()
        print(f"   {k:22s} rel-L2 {((gg - g).norm() / g.norm().clamp_min(1e-12)).item():.3e}  |g| {g.norm().item():.3e} |ours| {gg.norm().item():.3e}")
