"""Generate golden vectors by running the UNMODIFIED reference modules on CPU.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

The reference has no tests or fixtures of its own, so these vectors -- outputs
of the reference's own functions on seeded inputs -- are what pins oracle/oracle.py and, through
it, the CUDA path.  The .npz files written next to this script are committed; nothing at test
time reads /root/reference.
"""
import importlib.util
import os
import sys

import numpy as np
import torch

REF = os.environ.get("TNERF_REFERENCE", "/root/reference/src")
HERE = os.path.dirname(os.path.abspath(__file__))


def ref_module(name):
    spec = importlib.util.spec_from_file_location(f"_ref_{name}", os.path.join(REF, f"{name}.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def pose(theta, phi, r=4.0):
    sys.path.insert(0, os.path.join(HERE, "..", ".."))
    from oracle.oracle import look_at_pose
    return look_at_pose(theta, phi, r)


def main():
    torch.set_num_threads(1)
    rays, sampling, encoding, nerf, volume, utils, camera = (
        ref_module(n) for n in ("rays", "sampling", "encoding", "nerf", "volume", "utils", "camera"))
    out = {}

    # ---- a1 get_rays: odd sizes, the dataset's focal, an arbitrary pose
    focal = float(np.float32(138.88888549804688))
    c2w = pose(0.7, 0.45)
    for tag, (H, W) in {"small": (5, 7), "tile": (12, 10)}.items():
        ro, rd = rays.get_rays(H, W, focal, c2w)
        out[f"rays_{tag}_HWf"] = np.array([H, W, focal], dtype=np.float64)
        out[f"rays_{tag}_c2w"] = c2w.numpy()
        out[f"rays_{tag}_o"] = ro.contiguous().numpy()
        out[f"rays_{tag}_d"] = rd.numpy()
    ro, rd = rays.get_rays(100, 100, focal, c2w)
    pick = torch.tensor([0, 1, 99, 100, 4999, 5050, 9899, 9999])
    out["rays_full_pick"] = pick.numpy()
    out["rays_full_d"] = rd[pick].numpy()

    # ---- a3 stratified: deterministic + jittered (jitter reproduced as torch.rand(N,S) under a seed)
    ro, rd = rays.get_rays(5, 7, focal, c2w)
    for S in (8, 64):
        z, pts = sampling.stratified_samples(2.0, 6.0, S, ro, rd, randomized=False)
        out[f"strat_det_S{S}_z"] = z.contiguous().numpy()
        out[f"strat_det_S{S}_pts"] = pts.numpy()
        torch.manual_seed(1234 + S)
        z, pts = sampling.stratified_samples(2.0, 6.0, S, ro, rd, randomized=True)
        torch.manual_seed(1234 + S)
        u = torch.rand(ro.shape[0], S)
        out[f"strat_rand_S{S}_u"] = u.numpy()
        out[f"strat_rand_S{S}_z"] = z.numpy()
        out[f"strat_rand_S{S}_pts"] = pts.numpy()
    out["strat_ro"] = ro.contiguous().numpy()
    out["strat_rd"] = rd.numpy()
    # near/far as per-ray tensors (docstring of sampling.py:8)
    nr = torch.linspace(1.5, 2.5, ro.shape[0]).unsqueeze(1)
    fr = torch.linspace(5.0, 7.0, ro.shape[0]).unsqueeze(1)
    z, pts = sampling.stratified_samples(nr, fr, 8, ro, rd, randomized=False)
    out["strat_tensor_near"] = nr.numpy(); out["strat_tensor_far"] = fr.numpy()
    out["strat_tensor_z"] = z.numpy()

    # ---- a4 positional encoding, L in {2,6,10}, with and without the raw input
    g = torch.Generator().manual_seed(7)
    x = (torch.rand(33, 3, generator=g) * 2 - 1) * 6.5
    out["enc_x"] = x.numpy()
    for L in (2, 6, 10):
        for inc in (True, False):
            e = encoding.PositionalEncoding(num_freqs=L, include_input=inc)
            out[f"enc_L{L}_{int(inc)}"] = e(x).numpy()

    # ---- a5 TinyNeRF forward: repo MLP (63,128,4,2) on 16 rows; small MLPs for other skip positions
    torch.manual_seed(0)
    e10 = encoding.PositionalEncoding(10, True)
    feat = e10(x[:16])
    net = nerf.TinyNeRF(63, 128, 4, 2)
    for k, v in net.state_dict().items():
        out[f"mlp_repo_p_{k}"] = v.numpy()
    with torch.no_grad():
        c, s = net(feat)
    out["mlp_repo_x"] = feat.numpy(); out["mlp_repo_rgb"] = c.numpy(); out["mlp_repo_sigma"] = s.numpy()
    for tag, (ind, hid, dep, sk) in {"a": (39, 32, 4, 2), "b": (15, 16, 3, 1), "c": (15, 16, 2, 0), "d": (39, 32, 5, 4)}.items():
        torch.manual_seed(11)
        L = (ind - 3) // 6
        net = nerf.TinyNeRF(ind, hid, dep, sk)
        f = encoding.PositionalEncoding(L, True)(x[:9])
        with torch.no_grad():
            c, s = net(f)
        out[f"mlp_{tag}_cfg"] = np.array([ind, hid, dep, sk])
        for k, v in net.state_dict().items():
            out[f"mlp_{tag}_p_{k}"] = v.numpy()
        out[f"mlp_{tag}_x"] = f.numpy(); out[f"mlp_{tag}_rgb"] = c.numpy(); out[f"mlp_{tag}_sigma"] = s.numpy()

    # ---- a6 volume_render: random, all-zero density, opaque, scaled ray dirs, black background
    g = torch.Generator().manual_seed(3)
    N, S = 6, 16
    rgb = torch.rand(N, S, 3, generator=g)
    sig = torch.relu(torch.randn(N, S, 1, generator=g) * 3.0)
    z = torch.sort(torch.rand(N, S, generator=g) * 4 + 2, dim=1).values
    rdir = torch.randn(N, 3, generator=g) * 1.7
    sig[1] = 0.0
    sig[2] = 50.0
    sig[3, -1] = 0.0
    out["vol_rgb"], out["vol_sigma"], out["vol_z"], out["vol_rd"] = rgb.numpy(), sig.numpy(), z.numpy(), rdir.numpy()
    for wb in (True, False):
        a = rgb.clone().requires_grad_(True)
        b = sig.clone().requires_grad_(True)
        c, d, acc, w = volume.volume_render(a, b, z, rdir, white_bkgd=wb)
        gC = torch.rand(N, 3, generator=torch.Generator().manual_seed(5)) - 0.5
        gD = torch.rand(N, 1, generator=torch.Generator().manual_seed(6)) - 0.5
        gA = torch.rand(N, 1, generator=torch.Generator().manual_seed(8)) - 0.5
        gW = torch.rand(N, S, generator=torch.Generator().manual_seed(9)) - 0.5
        ((c * gC).sum() + (d * gD).sum() + (acc * gA).sum() + (w * gW).sum()).backward()
        t = int(wb)
        out[f"vol{t}_c"], out[f"vol{t}_d"], out[f"vol{t}_a"], out[f"vol{t}_w"] = (
            c.detach().numpy(), d.detach().numpy(), acc.detach().numpy(), w.detach().numpy())
        out[f"vol{t}_grgb"], out[f"vol{t}_gsigma"] = a.grad.numpy(), b.grad.numpy()
        out["vol_gC"], out["vol_gD"], out["vol_gA"], out["vol_gW"] = gC.numpy(), gD.numpy(), gA.numpy(), gW.numpy()

    # ---- a7 psnr
    m = torch.tensor([1e-12, 1e-3, 0.07, 1.0])
    out["psnr_in"] = m.numpy(); out["psnr_out"] = utils.mse2psnr(m).numpy()

    # ---- a2..a9 composed: three train steps of the reference loop body on CPU (train.py:108-128),
    #      small MLP so the fixture stays small; full repo MLP covered by mlp_repo above.
    torch.manual_seed(0)
    enc = encoding.PositionalEncoding(4, True)
    net = nerf.TinyNeRF(enc.out_dim, 32, 4, 2)
    opt = torch.optim.Adam(net.parameters(), lr=5e-4)
    for k, v in net.state_dict().items():
        out[f"train_p0_{k}"] = v.numpy().copy()
    ro_all, rd_all = rays.get_rays(12, 10, focal, c2w)
    gi = torch.Generator().manual_seed(21)
    pix = torch.rand(120, 3, generator=gi)
    for step in range(3):
        inds = torch.randint(0, 120, (24,), generator=gi)
        ro, rd, tgt = ro_all[inds], rd_all[inds], pix[inds]
        torch.manual_seed(500 + step)
        z, pts = sampling.stratified_samples(2.0, 6.0, 16, ro, rd, randomized=True)
        torch.manual_seed(500 + step)
        u = torch.rand(24, 16)
        c, s = net(enc(pts.reshape(-1, 3)))
        comp, dep, acc, _ = volume.volume_render(c.reshape(24, 16, 3), s.reshape(24, 16, 1), z, rd)
        loss = torch.mean((comp - tgt) ** 2)
        opt.zero_grad(set_to_none=True)
        loss.backward()
        out[f"train_s{step}_inds"] = inds.numpy(); out[f"train_s{step}_u"] = u.numpy()
        out[f"train_s{step}_loss"] = loss.detach().numpy()
        out[f"train_s{step}_comp"] = comp.detach().numpy()
        out[f"train_s{step}_depth"] = dep.detach().numpy()
        out[f"train_s{step}_acc"] = acc.detach().numpy()
        for k, v in net.named_parameters():
            out[f"train_s{step}_g_{k}"] = v.grad.numpy().copy()
        opt.step()
        for k, v in net.state_dict().items():
            out[f"train_s{step}_p_{k}"] = v.numpy().copy()
    out["train_c2w"] = c2w.numpy(); out["train_pix"] = pix.numpy()
    out["train_ro_all"] = ro_all.contiguous().numpy(); out["train_rd_all"] = rd_all.numpy()

    # ---- render_one equivalent (train.py:36-59) at 12x10, chunk 50 -> 3 chunks, reference modules only
    with torch.no_grad():
        chunks = []
        for a in range(0, 120, 50):
            z, pts = sampling.stratified_samples(2.0, 6.0, 16, ro_all[a:a + 50], rd_all[a:a + 50], randomized=False)
            c, s = net(enc(pts.reshape(-1, 3)))
            n = pts.shape[0]
            chunks.append(volume.volume_render(c.reshape(n, 16, 3), s.reshape(n, 16, 1), z, rd_all[a:a + 50])[0])
        out["render_img"] = torch.cat(chunks, 0).reshape(12, 10, 3).clamp(0, 1).numpy()

    # ---- N3 spiral poses
    out["spiral"] = camera.spiral_poses(c2w, 7, 0.3).numpy()

    path = os.path.join(HERE, "reference_vectors.npz")
    np.savez_compressed(path, **out)
    print(f"wrote {path}: {len(out)} arrays, {os.path.getsize(path) / 1024:.1f} KiB, torch {torch.__version__}")


if __name__ == "__main__":
    main()
