"""GPU parity tests of the stand-alone (unfused) ops: CUDA path through the C ABI vs the CPU oracle
and the committed reference vectors.  Tolerances follow BASELINE.json's north_star:
rays / samples 1e-6 relative, per-ray rgb/depth/acc 2e-3 absolute, gradients 1e-2 relative."""
import numpy as np
import pytest
import torch

from oracle import oracle as O

pytestmark = pytest.mark.gpu
T = torch.from_numpy


@pytest.fixture(scope="module")
def dev():
    return torch.device("cuda:0")


def rel_l2(a, b):
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def norm_close(a, b, tol=1e-6):
    """norm-wise relative agreement (components of unit vectors pass through 0)"""
    a, b = a.detach().cpu().double(), torch.as_tensor(b).double()
    scale = b.abs().max().clamp_min(1e-30)
    assert ((a - b).abs().max() / scale).item() <= tol, ((a - b).abs().max() / scale).item()


# ---------------------------------------------------------------------------------------------- a1
@pytest.mark.parametrize("tag", ["small", "tile"])
def test_get_rays_golden(golden, dev, tag):
    from rays import get_rays
    H, W, focal = golden[f"rays_{tag}_HWf"]
    ro, rd = get_rays(int(H), int(W), float(focal), T(golden[f"rays_{tag}_c2w"]).to(dev))
    assert ro.shape == rd.shape == (int(H) * int(W), 3) and ro.stride(0) == 0 and rd.dtype == torch.float32
    norm_close(ro, golden[f"rays_{tag}_o"], 0.0)
    norm_close(rd, golden[f"rays_{tag}_d"], 1e-6)


def test_get_rays_full_frame_vs_oracle(dev):
    from rays import get_rays
    pose = O.look_at_pose(1.1, 0.4)
    for (H, W, f) in ((100, 100, 138.88888549804688), (37, 53, 60.0), (800, 800, 1111.11)):
        ro, rd = get_rays(H, W, f, pose.to(dev))
        oro, ord_ = O.get_rays(H, W, f, pose)
        norm_close(rd, ord_, 1e-6)
        norm_close(ro, oro, 0.0)
        assert torch.allclose(rd.norm(dim=1), torch.ones(H * W, device=dev), atol=1e-6)


def test_get_rays_device_kw_and_cpu_pose(dev):
    from rays import get_rays
    pose = O.look_at_pose(0.2, 0.3)
    ro, rd = get_rays(4, 6, 10.0, pose, device=dev)     # CPU pose + device= (train.py:46 pattern)
    assert rd.is_cuda and ro.is_cuda
    with pytest.raises(RuntimeError):
        get_rays(4, 6, 10.0, pose)                        # no CPU path


# ---------------------------------------------------------------------------------------------- a3
@pytest.mark.parametrize("S", [8, 64])
def test_stratified_golden(golden, dev, S):
    from sampling import stratified_samples
    ro, rd = T(golden["strat_ro"]).to(dev), T(golden["strat_rd"]).to(dev)
    z, pts = stratified_samples(2.0, 6.0, S, ro, rd, randomized=False)
    assert z.shape == (ro.shape[0], S) and pts.shape == (ro.shape[0], S, 3)
    assert np.array_equal(z.cpu().numpy(), golden[f"strat_det_S{S}_z"])          # bit exact depths
    norm_close(pts + 0, golden[f"strat_det_S{S}_pts"], 1e-6)
    u = T(golden[f"strat_rand_S{S}_u"]).to(dev)
    z, pts = stratified_samples(2.0, 6.0, S, ro, rd, randomized=True, t_rand=u)
    assert np.array_equal(z.cpu().numpy(), golden[f"strat_rand_S{S}_z"])
    norm_close(pts + 0, golden[f"strat_rand_S{S}_pts"], 1e-6)


@pytest.mark.parametrize("S", [1, 2, 3, 33, 64, 128, 192, 257])
def test_stratified_bit_exact_vs_oracle(dev, S):
    from sampling import stratified_samples
    g = torch.Generator().manual_seed(S)
    n = 19
    ro = torch.randn(n, 3, generator=g)
    rd = torch.nn.functional.normalize(torch.randn(n, 3, generator=g), dim=1)
    u = torch.rand(n, S, generator=g)
    for near, far in ((2.0, 6.0), (0.05, 1.3)):
        z, pts = stratified_samples(near, far, S, ro.to(dev), rd.to(dev), randomized=True, t_rand=u.to(dev))
        oz, op = O.stratified(near, far, S, ro, rd, u)
        assert torch.equal(z.cpu(), oz)
        assert torch.equal((pts + 0).cpu(), op)
        z, pts = stratified_samples(near, far, S, ro.to(dev), rd.to(dev), randomized=False)
        oz, op = O.stratified(near, far, S, ro, rd, None)
        assert z.stride(0) == 0 and torch.equal(z.cpu(), oz) and torch.equal((pts + 0).cpu(), op)


def test_stratified_tensor_near_far_and_global_rng(golden, dev):
    from sampling import stratified_samples
    ro, rd = T(golden["strat_ro"]).to(dev), T(golden["strat_rd"]).to(dev)
    z, _ = stratified_samples(T(golden["strat_tensor_near"]).to(dev), T(golden["strat_tensor_far"]).to(dev), 8, ro, rd, randomized=False)
    assert np.array_equal(z.cpu().numpy(), golden["strat_tensor_z"])
    torch.manual_seed(5)
    z1, _ = stratified_samples(2.0, 6.0, 16, ro, rd, randomized=True)
    torch.manual_seed(5)
    z2, _ = stratified_samples(2.0, 6.0, 16, ro, rd, randomized=True)
    det = O.depth_bins(2.0, 6.0, 16)
    assert torch.equal(z1, z2) and not torch.equal(z1.cpu()[0], det)
    half = (det[1] - det[0]) / 2
    assert (z1.cpu() >= det - half - 1e-6).all() and (z1.cpu() <= det + half + 1e-6).all()


def test_stratified_empty(dev):
    from sampling import stratified_samples
    z, pts = stratified_samples(2.0, 6.0, 8, torch.zeros(0, 3, device=dev), torch.zeros(0, 3, device=dev), randomized=True)
    assert z.shape == (0, 8) and pts.shape == (0, 8, 3)


# ---------------------------------------------------------------------------------------------- a4
@pytest.mark.parametrize("L", [2, 6, 10])
@pytest.mark.parametrize("inc", [True, False])
def test_posenc_golden(golden, dev, L, inc):
    from encoding import PositionalEncoding
    enc = PositionalEncoding(L, inc).to(dev)
    assert enc.out_dim == 6 * L + (3 if inc else 0)
    assert list(enc.state_dict().keys()) == ["freq_bands"] and torch.equal(enc.freq_bands.cpu(), 2.0 ** torch.arange(L).float())
    out = enc(T(golden["enc_x"]).to(dev))
    ref = golden[f"enc_L{L}_{int(inc)}"]
    assert out.shape == ref.shape
    np.testing.assert_allclose(out.cpu().numpy(), ref, rtol=0, atol=3e-7)


def test_posenc_shapes_grad_and_assert(dev):
    from encoding import PositionalEncoding
    enc = PositionalEncoding(4, True).to(dev)
    x = (torch.rand(5, 7, 3, generator=torch.Generator().manual_seed(0)) * 4 - 2)
    xg = x.to(dev).requires_grad_(True)
    out = enc(xg)
    assert out.shape == (5, 7, 27)
    w = torch.randn(5, 7, 27, generator=torch.Generator().manual_seed(1))
    (out * w.to(dev)).sum().backward()
    xr = x.clone().requires_grad_(True)
    (O.posenc(xr, 4, True) * w).sum().backward()
    assert rel_l2(xg.grad.cpu(), xr.grad) < 1e-5
    with pytest.raises(AssertionError):
        enc(torch.zeros(3, 2, device=dev))
    assert enc(torch.zeros(0, 3, device=dev)).shape == (0, 27)


# ---------------------------------------------------------------------------------------------- a5
def load_model(golden, prefix, cfg, dev):
    from nerf import TinyNeRF
    m = TinyNeRF(*cfg)
    sd = {k[len(prefix):]: T(golden[k].copy()) for k in golden.files if k.startswith(prefix)}
    m.load_state_dict(sd)
    return m.to(dev), sd


def test_mlp_repo_golden(golden, dev):
    m, sd = load_model(golden, "mlp_repo_p_", (63, 128, 4, 2), dev)
    assert [k for k, _ in m.state_dict().items()] == [k for k, _ in O.mlp_param_shapes(63, 128, 4, 2)]
    rgb, sigma = m(T(golden["mlp_repo_x"]).to(dev))
    assert rgb.shape == (16, 3) and sigma.shape == (16, 1)
    np.testing.assert_allclose(rgb.detach().cpu().numpy(), golden["mlp_repo_rgb"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(sigma.detach().cpu().numpy(), golden["mlp_repo_sigma"], rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("tag", ["a", "b", "c", "d"])
def test_mlp_variants_golden(golden, dev, tag):
    cfg = tuple(int(v) for v in golden[f"mlp_{tag}_cfg"])
    m, _ = load_model(golden, f"mlp_{tag}_p_", cfg, dev)
    rgb, sigma = m(T(golden[f"mlp_{tag}_x"]).to(dev))
    np.testing.assert_allclose(rgb.detach().cpu().numpy(), golden[f"mlp_{tag}_rgb"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(sigma.detach().cpu().numpy(), golden[f"mlp_{tag}_sigma"], rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("cfg", [(63, 128, 4, 2), (39, 64, 3, 1), (27, 48, 5, 0), (15, 32, 2, 1)])
def test_mlp_forward_backward_vs_oracle(dev, cfg):
    from nerf import TinyNeRF
    ind, hid, dep, sk = cfg
    p = O.init_params(ind, hid, dep, sk, seed=3)
    m = TinyNeRF(ind, hid, dep, sk)
    m.load_state_dict(p)
    m = m.to(dev)
    g = torch.Generator().manual_seed(4)
    n = 777
    x = torch.rand(n, ind, generator=g) * 2 - 1
    wr, ws = torch.randn(n, 3, generator=g), torch.randn(n, 1, generator=g)
    xg = x.to(dev).requires_grad_(True)
    rgb, sigma = m(xg)
    ((rgb * wr.to(dev)).sum() + (sigma * ws.to(dev)).sum()).backward()
    q = {k: v.clone().requires_grad_(True) for k, v in p.items()}
    xr = x.clone().requires_grad_(True)
    orgb, osig = O.mlp_forward(q, xr, dep, sk)
    ((orgb * wr).sum() + (osig * ws).sum()).backward()
    assert (rgb.detach().cpu() - orgb.detach()).abs().max() < 1e-5
    assert (sigma.detach().cpu() - osig.detach()).abs().max() < 1e-5
    for k, v in m.named_parameters():
        assert rel_l2(v.grad.cpu(), q[k].grad) < 1e-4, k
    assert rel_l2(xg.grad.cpu(), xr.grad) < 1e-4


def test_mlp_skip_equal_depth_errors_like_reference(dev):
    from nerf import TinyNeRF
    m = TinyNeRF(15, 16, 2, 2).to(dev)
    with pytest.raises(RuntimeError):
        m(torch.zeros(4, 15, device=dev))


def test_mlp_no_grad_and_empty(dev):
    from nerf import TinyNeRF
    m = TinyNeRF(15, 16, 2, 1).to(dev)
    with torch.no_grad():
        rgb, sigma = m(torch.zeros(0, 15, device=dev))
    assert rgb.shape == (0, 3) and sigma.shape == (0, 1)


# ---------------------------------------------------------------------------------------------- a6
@pytest.mark.parametrize("wb", [1, 0])
def test_volume_render_golden(golden, dev, wb):
    from volume import volume_render
    rgb, sig, z, rd = (T(golden[k]).to(dev) for k in ("vol_rgb", "vol_sigma", "vol_z", "vol_rd"))
    rgb.requires_grad_(True); sig.requires_grad_(True)
    c, d, a, w = volume_render(rgb, sig, z, rd, white_bkgd=bool(wb))
    assert c.shape == (6, 3) and d.shape == (6, 1) and a.shape == (6, 1) and w.shape == (6, 16)
    for got, key in ((c, "c"), (d, "d"), (a, "a"), (w, "w")):
        np.testing.assert_allclose(got.detach().cpu().numpy(), golden[f"vol{wb}_{key}"], rtol=2e-6, atol=2e-7)
    gC, gD, gA, gW = (T(golden[k]).to(dev) for k in ("vol_gC", "vol_gD", "vol_gA", "vol_gW"))
    ((c * gC).sum() + (d * gD).sum() + (a * gA).sum() + (w * gW).sum()).backward()
    np.testing.assert_allclose(rgb.grad.cpu().numpy(), golden[f"vol{wb}_grgb"], rtol=1e-5, atol=1e-7)
    ref = golden[f"vol{wb}_gsigma"]
    np.testing.assert_allclose(sig.grad.cpu().numpy(), ref, rtol=2e-4, atol=1e-5 * np.abs(ref).max())


@pytest.mark.parametrize("S", [1, 5, 32, 64, 100, 192, 256])
def test_volume_render_vs_oracle(dev, S):
    from volume import volume_render
    g = torch.Generator().manual_seed(S)
    n = 41
    rgb = torch.rand(n, S, 3, generator=g)
    sig = torch.relu(torch.randn(n, S, 1, generator=g) * 4)
    z = torch.sort(torch.rand(n, S, generator=g) * 4 + 2, dim=1).values
    rd = torch.randn(n, 3, generator=g)
    gC, gD, gA = torch.randn(n, 3, generator=g), torch.randn(n, 1, generator=g), torch.randn(n, 1, generator=g)
    a = rgb.to(dev).requires_grad_(True); b = sig.to(dev).requires_grad_(True)
    c, d, acc, w = volume_render(a, b, z.to(dev), rd.to(dev))
    ((c * gC.to(dev)).sum() + (d * gD.to(dev)).sum() + (acc * gA.to(dev)).sum()).backward()
    oc, od, oa, ow = O.composite(rgb, sig, z, rd, True)
    for got, ref in ((c, oc), (d, od), (acc, oa), (w, ow)):
        assert (got.detach().cpu() - ref).abs().max() < 2e-6 * max(1.0, ref.abs().max().item())
    dr, ds = O.composite_backward(rgb.double(), sig.double(), z.double(), rd.double(), gC.double(), gD.double(), gA.double(), None, True)
    assert rel_l2(a.grad.cpu().double(), dr) < 1e-5
    assert rel_l2(b.grad.cpu().double(), ds) < 1e-4


def test_volume_render_edge_cases(dev):
    from volume import volume_render
    n, S = 3, 64
    z = O.depth_bins(2.0, 6.0, S).expand(n, S).to(dev)
    rd = torch.tensor([[0.0, 0.0, -1.0]] * n, device=dev)
    rgb = torch.full((n, S, 3), 0.25, device=dev)
    c, d, a, w = volume_render(rgb, torch.zeros(n, S, 1, device=dev), z, rd)                     # empty space
    assert torch.allclose(c, torch.ones_like(c)) and torch.all(a == 0) and torch.all(d == 0)
    c, d, a, w = volume_render(rgb, torch.full((n, S, 1), 1e4, device=dev), z, rd)               # opaque wall
    assert torch.allclose(a, torch.ones_like(a)) and torch.allclose(w[:, 0], torch.ones(n, device=dev))
    assert torch.allclose(w[:, 1], torch.full((n,), 1e-10, device=dev), rtol=1e-3, atol=0)
    sig = torch.zeros(n, S, 1, device=dev); sig[:, -1] = 1e-6                                     # only the 1e10 tail sample
    c, d, a, w = volume_render(rgb, sig, z, rd)
    assert torch.allclose(a, torch.ones_like(a))
    c2, *_ = volume_render(rgb, sig, z, rd, white_bkgd=False)
    assert torch.allclose(c2, torch.full_like(c2, 0.25), atol=1e-6)
    # fp16 inputs under autocast are promoted (reference: volume.py:23 deltas are fp32)
    with torch.autocast("cuda"):
        c3, *_ = volume_render(rgb.half(), sig.half(), z, rd)
    assert c3.dtype == torch.float32


# ---------------------------------------------------------------------------------------------- a7, a9
def test_psnr_and_mse_kernel(golden, dev):
    import _engine as E
    from utils import mse2psnr
    np.testing.assert_allclose(mse2psnr(T(golden["psnr_in"]).to(dev)).cpu().numpy(), golden["psnr_out"], rtol=1e-6)
    g = torch.Generator().manual_seed(0)
    a, b = torch.rand(4096, 3, generator=g), torch.rand(4096, 3, generator=g)
    out = torch.empty(2, device=dev)
    a_d, b_d = a.to(dev), b.to(dev)
    E.check(E.lib().tnerf_mse_psnr(E.ptr(a_d), E.ptr(b_d), a.numel(), E.ptr(out), E.stream(dev)))
    m = O.mse(a, b)
    assert abs(out[0].item() - m.item()) < 1e-6 and abs(out[1].item() - O.mse2psnr(m).item()) < 1e-4


def test_adam_kernel_matches_torch_optim(dev):
    import _engine as E
    g = torch.Generator().manual_seed(0)
    n = 66308
    p0 = torch.randn(n, generator=g) * 0.1
    p_ref = p0.clone().requires_grad_(True)
    opt = torch.optim.Adam([p_ref], lr=5e-4)
    p, m, v = p0.to(dev), torch.zeros(n, device=dev), torch.zeros(n, device=dev)
    flag = torch.zeros(1, dtype=torch.int32, device=dev)
    for step in range(1, 6):
        grad = torch.randn(n, generator=g) * (10.0 ** (step - 4))
        p_ref.grad = grad.clone()
        opt.step()
        g_d, g8_d = grad.to(dev), (grad * 8).to(dev)
        E.check(E.lib().tnerf_check_finite(E.ptr(g_d), n, E.ptr(flag), E.stream(dev)))
        E.check(E.lib().tnerf_adam_step(E.ptr(p), E.ptr(g8_d), E.ptr(m), E.ptr(v), n, step, 5e-4, 0.9, 0.999, 1e-8,
                                        1.0 / 8, E.ptr(flag), E.stream(dev)))
        assert flag.item() == 0
        assert (p.cpu() - p_ref.detach()).abs().max() < 2e-7
    bad = torch.randn(n, generator=g); bad[123] = float("inf")
    before = p.clone()
    bad_d = bad.to(dev)
    E.check(E.lib().tnerf_check_finite(E.ptr(bad_d), n, E.ptr(flag), E.stream(dev)))
    E.check(E.lib().tnerf_adam_step(E.ptr(p), E.ptr(bad_d), E.ptr(m), E.ptr(v), n, 6, 5e-4, 0.9, 0.999, 1e-8, 1.0, E.ptr(flag), E.stream(dev)))
    assert flag.item() == 1 and torch.equal(p, before)          # GradScaler semantics: skipped step


def test_gather3(dev):
    import _engine as E
    g = torch.Generator().manual_seed(0)
    a, b, c = (torch.randn(1000, 3, generator=g).to(dev) for _ in range(3))
    idx = torch.randint(0, 1000, (257,), generator=g).to(dev)
    oa, ob, oc = (torch.empty(257, 3, device=dev) for _ in range(3))
    E.check(E.lib().tnerf_gather3(E.ptr(idx), 257, 1000, E.ptr(a), E.ptr(oa), E.ptr(b), E.ptr(ob), E.ptr(c), E.ptr(oc), E.stream(dev)))
    assert torch.equal(oa, a[idx]) and torch.equal(ob, b[idx]) and torch.equal(oc, c[idx])
