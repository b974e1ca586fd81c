"""Golden vectors for BASELINE config 4 (hidden 256, 800x800 frame, 192 samples/ray), produced by the UNMODIFIED reference
modules on CPU.  Run in the build container only (needs /root/reference):

    python tests/golden/make_golden_c4.py

The hidden-256 parameters (459 k floats) are not stored: both this script and the test derive them from
`synthetic_params` below (seeded torch CPU generator), so the fixture holds only inputs and the reference's outputs.
"""
import math
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(HERE, "..", ".."))


def synthetic_params(in_dim=63, hidden=256, depth=4, skip_at=2, seed=4321, gain=1.5):
    """state_dict of TinyNeRF(in_dim, hidden, depth, skip_at) (src/nerf.py:10-27 key order) from a seeded CPU generator"""
    g = torch.Generator().manual_seed(seed)
    p = {}

    def lin(prefix, out, fan):
        b = 1.0 / math.sqrt(fan)
        p[f"{prefix}.weight"] = (torch.rand(out, fan, generator=g) * 2 - 1) * b * gain
        p[f"{prefix}.bias"] = (torch.rand(out, generator=g) * 2 - 1) * b
    for l in range(depth):
        lin(f"layers.{l}", hidden, in_dim if l == 0 else (hidden + in_dim if l == skip_at else hidden))
    lin("sigma.0", 1, hidden)
    lin("rgb.0", 3, hidden)
    p["sigma.0.bias"] = p["sigma.0.bias"] + 0.3
    return p


def main():
    from make_golden import ref_module, pose
    torch.set_num_threads(1)
    rays, sampling, encoding, nerf, volume = (ref_module(n) for n in ("rays", "sampling", "encoding", "nerf", "volume"))
    out = {}
    net = nerf.TinyNeRF(63, hidden=256)
    net.load_state_dict(synthetic_params())
    enc = encoding.PositionalEncoding(10, True)
    c2w = pose(0.4, 0.55)
    H = W = 800
    focal = 1111.11
    ro, rd = rays.get_rays(H, W, focal, c2w)
    pick = torch.tensor([0, 799, 320400, 320401, 333333, 400 * 800 + 17, 639999, 123456, 500000, 77, 250250, 600600])
    ro_p, rd_p = ro[pick].contiguous(), rd[pick].contiguous()
    with torch.no_grad():
        z, pts = sampling.stratified_samples(2.0, 6.0, 192, ro_p, rd_p, randomized=False)
        feat = enc(pts.reshape(-1, 3))
        rgb, sigma = net(feat)
        comp, depth, acc, w = volume.volume_render(rgb.reshape(-1, 192, 3), sigma.reshape(-1, 192, 1), z, rd_p)
    out["c4_c2w"] = c2w.numpy(); out["c4_pick"] = pick.numpy()
    out["c4_rays_d"] = rd_p.numpy(); out["c4_rays_o"] = ro_p.numpy()
    out["c4_z"] = np.ascontiguousarray(z.numpy()); out["c4_feat_rows"] = feat[::97].numpy()
    out["c4_rgb"] = rgb.numpy(); out["c4_sigma"] = sigma.numpy()
    out["c4_comp"] = comp.numpy(); out["c4_depth"] = depth.numpy(); out["c4_acc"] = acc.numpy(); out["c4_weights"] = w.numpy()
    path = os.path.join(HERE, "reference_vectors_c4.npz")
    np.savez_compressed(path, **out)
    print(f"wrote {path}: {len(out)} arrays, {os.path.getsize(path)} bytes")


if __name__ == "__main__":
    main()
