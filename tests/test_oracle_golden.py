"""Pin oracle/oracle.py against vectors produced by the unmodified reference
(tests/golden/make_golden.py).  CPU only.  Element-wise ops must match bit for bit when run with
the torch build that generated them; GEMM-bearing paths get a few-ulp tolerance because MKL's
blocking (and therefore summation order) depends on the host CPU."""
import numpy as np
import pytest
import torch

from oracle import oracle as O

T = torch.from_numpy


def close(a, b, rtol=0.0, atol=0.0):
    a = a.detach().numpy() if isinstance(a, torch.Tensor) else a
    np.testing.assert_allclose(a, b, rtol=rtol, atol=atol)


def params_from(golden, prefix):
    return {k[len(prefix):]: T(golden[k].copy()) for k in golden.files if k.startswith(prefix)}


@pytest.mark.parametrize("tag", ["small", "tile"])
def test_get_rays(golden, tag):
    H, W, focal = golden[f"rays_{tag}_HWf"]
    ro, rd = O.get_rays(int(H), int(W), float(focal), T(golden[f"rays_{tag}_c2w"]))
    close(ro, golden[f"rays_{tag}_o"])
    close(rd, golden[f"rays_{tag}_d"], rtol=1e-6, atol=1e-7)


def test_get_rays_full_frame_rows(golden):
    _, rd = O.get_rays(100, 100, float(np.float32(138.88888549804688)), T(golden["rays_small_c2w"]))
    close(rd[T(golden["rays_full_pick"])], golden["rays_full_d"], rtol=1e-6, atol=1e-7)


@pytest.mark.parametrize("S", [8, 64])
def test_stratified(golden, S):
    ro, rd = T(golden["strat_ro"]), T(golden["strat_rd"])
    z, pts = O.stratified(2.0, 6.0, S, ro, rd, None)
    close(z, golden[f"strat_det_S{S}_z"])
    close(pts, golden[f"strat_det_S{S}_pts"])
    z, pts = O.stratified(2.0, 6.0, S, ro, rd, T(golden[f"strat_rand_S{S}_u"]))
    close(z, golden[f"strat_rand_S{S}_z"])
    close(pts, golden[f"strat_rand_S{S}_pts"])


def test_stratified_first_last_bins_half_width(golden):
    z = golden["strat_det_S8_z"][0]
    zr = golden["strat_rand_S8_z"]
    step = z[1] - z[0]
    assert np.all(zr[:, 0] <= z[0] + step / 2 + 1e-6) and np.all(zr[:, 0] >= z[0])
    assert np.all(zr[:, -1] >= z[-1] - step / 2 - 1e-6) and np.all(zr[:, -1] <= z[-1])


def test_stratified_tensor_near_far(golden):
    ro, rd = T(golden["strat_ro"]), T(golden["strat_rd"])
    z, _ = O.stratified(T(golden["strat_tensor_near"]), T(golden["strat_tensor_far"]), 8, ro, rd, None)
    close(z, golden["strat_tensor_z"])


@pytest.mark.parametrize("L", [2, 6, 10])
@pytest.mark.parametrize("inc", [True, False])
def test_posenc(golden, L, inc):
    e = O.posenc(T(golden["enc_x"]), L, inc)
    assert e.shape[-1] == O.posenc_dim(L, inc)
    close(e, golden[f"enc_L{L}_{int(inc)}"])


def test_posenc_rejects_wrong_width():
    with pytest.raises(AssertionError):
        O.posenc(torch.zeros(4, 2))


def test_mlp_repo(golden):
    p = params_from(golden, "mlp_repo_p_")
    assert [tuple(v.shape) for v in p.values()] == [s for _, s in O.mlp_param_shapes(63, 128, 4, 2)]
    assert sum(v.numel() for v in p.values()) == 66308
    c, s = O.mlp_forward(p, T(golden["mlp_repo_x"]), 4, 2)
    close(c, golden["mlp_repo_rgb"], rtol=2e-6, atol=2e-7)
    close(s, golden["mlp_repo_sigma"], rtol=2e-6, atol=2e-6)


@pytest.mark.parametrize("tag", ["a", "b", "c", "d"])
def test_mlp_variants(golden, tag):
    ind, hid, dep, sk = (int(v) for v in golden[f"mlp_{tag}_cfg"])
    p = params_from(golden, f"mlp_{tag}_p_")
    assert list(p.keys()) == [k for k, _ in O.mlp_param_shapes(ind, hid, dep, sk)]
    c, s = O.mlp_forward(p, T(golden[f"mlp_{tag}_x"]), dep, sk)
    close(c, golden[f"mlp_{tag}_rgb"], rtol=2e-6, atol=2e-7)
    close(s, golden[f"mlp_{tag}_sigma"], rtol=2e-6, atol=2e-6)


@pytest.mark.parametrize("wb", [1, 0])
def test_composite_forward_backward(golden, wb):
    rgb, sig, z, rd = (T(golden[k]) for k in ("vol_rgb", "vol_sigma", "vol_z", "vol_rd"))
    c, d, a, w = O.composite(rgb, sig, z, rd, bool(wb))
    close(c, golden[f"vol{wb}_c"], rtol=1e-6, atol=1e-7)
    close(d, golden[f"vol{wb}_d"], rtol=1e-6, atol=1e-7)
    close(a, golden[f"vol{wb}_a"], rtol=1e-6, atol=1e-7)
    close(w, golden[f"vol{wb}_w"], rtol=1e-6, atol=1e-12)
    # edge rows: sigma == 0 -> white / zero acc; opaque -> acc == 1
    assert np.allclose(golden["vol1_a"][1], 0.0) and np.allclose(golden["vol1_c"][1], 1.0)
    assert np.allclose(golden["vol1_a"][2], 1.0, atol=1e-6)
    gC, gD, gA, gW = (T(golden[k]) for k in ("vol_gC", "vol_gD", "vol_gA", "vol_gW"))
    d_rgb, d_sig = O.composite_backward(rgb.double(), sig.double(), z.double(), rd.double(),
                                        gC.double(), gD.double(), gA.double(), gW.double(), bool(wb))
    close(d_rgb, golden[f"vol{wb}_grgb"], rtol=1e-5, atol=1e-7)
    # fp32 autograd of cumprod divides by q; compare where the reference's own value is well conditioned
    ref = golden[f"vol{wb}_gsigma"]
    np.testing.assert_allclose(d_sig.numpy(), ref, rtol=2e-4, atol=1e-5 * np.abs(ref).max())


def test_psnr(golden):
    close(O.mse2psnr(T(golden["psnr_in"])), golden["psnr_out"])


def test_three_train_steps(golden):
    """Reference loop body (train.py:108-128) replayed: loss, grads and Adam-updated parameters."""
    p = params_from(golden, "train_p0_")
    m = {k: torch.zeros_like(v) for k, v in p.items()}
    v = {k: torch.zeros_like(x) for k, x in p.items()}
    ro_all, rd_all, pix = T(golden["train_ro_all"]), T(golden["train_rd_all"]), T(golden["train_pix"])
    for step in range(3):
        inds = T(golden[f"train_s{step}_inds"])
        loss, g, (comp, dep, acc) = O.loss_and_grads(p, ro_all[inds], rd_all[inds], pix[inds], 2.0, 6.0, 16,
                                                     T(golden[f"train_s{step}_u"]), num_freqs=4)
        close(loss, golden[f"train_s{step}_loss"], rtol=1e-5)
        close(comp, golden[f"train_s{step}_comp"], rtol=1e-5, atol=1e-6)
        close(dep, golden[f"train_s{step}_depth"], rtol=1e-5, atol=1e-6)
        close(acc, golden[f"train_s{step}_acc"], rtol=1e-5, atol=1e-6)
        for k in p:
            ref = golden[f"train_s{step}_g_{k}"]
            np.testing.assert_allclose(g[k].numpy(), ref, rtol=1e-4, atol=1e-6 * max(1e-30, np.abs(ref).max()) + 1e-9)
        O.adam_step(p, g, m, v, step + 1)
        for k in p:
            close(p[k], golden[f"train_s{step}_p_{k}"], rtol=1e-5, atol=2e-6)


def test_render_image(golden):
    p = params_from(golden, "train_s2_p_")
    img = O.render_image(p, 12, 10, float(np.float32(138.88888549804688)), T(golden["train_c2w"]),
                         n_samples=16, chunk=50, num_freqs=4)
    close(img, golden["render_img"], rtol=1e-5, atol=1e-6)


def test_spiral(golden):
    close(O.spiral_poses(T(golden["train_c2w"]), 7, 0.3), golden["spiral"], rtol=1e-6, atol=1e-7)


# ------------------------------------------------------------------------------------------ BASELINE config 4 (hidden 256)
def c4_params():
    import importlib.util
    import os
    spec = importlib.util.spec_from_file_location("_make_golden_c4", os.path.join(os.path.dirname(__file__), "golden", "make_golden_c4.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.synthetic_params()


def test_config4_chain_against_reference(golden_c4):
    """800x800 frame, 192 deterministic samples, TinyNeRF(63, hidden=256): rays, depths, features, MLP outputs and the
    composited pixels of twelve rays, as the unmodified reference computes them."""
    g = golden_c4
    p = c4_params()
    ro, rd = O.get_rays(800, 800, 1111.11, T(g["c4_c2w"]))
    pick = T(g["c4_pick"])
    close(rd[pick], g["c4_rays_d"], rtol=1e-6, atol=1e-7)
    close(ro[pick], g["c4_rays_o"])
    ro_p, rd_p = T(g["c4_rays_o"].copy()), T(g["c4_rays_d"].copy())
    z, pts = O.stratified(2.0, 6.0, 192, ro_p, rd_p, None)
    close(z, g["c4_z"])
    feat = O.posenc(pts.reshape(-1, 3), 10, True)
    close(feat[::97], g["c4_feat_rows"], rtol=1e-6, atol=1e-6)
    c, s = O.mlp_forward(p, feat, 4, 2)
    close(c, g["c4_rgb"], rtol=1e-5, atol=2e-6)
    close(s, g["c4_sigma"], rtol=1e-5, atol=2e-5)
    comp, depth, acc, w = O.render_rays(p, ro_p, rd_p, 2.0, 6.0, 192, None)
    close(comp, g["c4_comp"], atol=2e-5)
    close(depth, g["c4_depth"], atol=1e-4)
    close(acc, g["c4_acc"], atol=2e-5)
    close(w, g["c4_weights"], atol=2e-5)


# ------------------------------------------------------------------------------------------ in-kernel jitter generator
def test_philox_known_answers():
    """The stratified jitter of the training kernel is Philox4x32-10 (third-party algorithm: Salmon et al., SC'11; not part of the
    reference, which calls torch.rand_like, src/sampling.py:24).  The oracle's restatement against the known-answer vectors published
    with the Random123 distribution (kat_vectors: counter, key -> output)."""
    kats = [((0, 0, 0, 0), (0, 0), (0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8)),
            ((0xFFFFFFFF,) * 4, (0xFFFFFFFF,) * 2, (0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD)),
            ((0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344), (0xA4093822, 0x299F31D0), (0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1))]
    for ctr, key, want in kats:
        assert [int(x) for x in O.philox4x32_10(ctr, key)] == list(want)
    # vectorised evaluation = element-wise evaluation
    c0 = np.arange(7, dtype=np.uint64)
    vec = O.philox4x32_10((c0, 5, 0, 9), (11, 13))[0]
    assert [int(O.philox4x32_10((int(i), 5, 0, 9), (11, 13))[0]) for i in c0] == [int(x) for x in vec]


def test_jitter_uniform_is_a_pure_function_of_seed_step_ray_sample():
    """u(seed, step, ray, sample): 24-bit uniform in [0, 1); a shard of the rays sees the numbers of the same rays in the full batch
    (ray-sharded data parallel draws per global ray id only through the per-rank seed -- the function itself is geometry-free)."""
    u = O.jitter_uniform(0x1234567887654321, (1 << 32) + 5, 96, 64)
    assert u.dtype == torch.float32 and u.shape == (96, 64)
    assert 0.0 <= float(u.min()) and float(u.max()) < 1.0
    assert torch.equal(u * 2 ** 24, (u * 2 ** 24).round())                 # 24-bit grid
    assert torch.equal(O.jitter_uniform(0x1234567887654321, (1 << 32) + 5, 40, 64, first_ray=56), u[56:])
    assert not torch.equal(O.jitter_uniform(0x1234567887654321, (1 << 32) + 6, 96, 64), u)      # next step: new numbers
    assert not torch.equal(O.jitter_uniform(0x1234567887654322, (1 << 32) + 5, 96, 64), u)      # other seed (rank): new numbers
    big = O.jitter_uniform(7, 1, 4096, 64)
    assert abs(float(big.mean()) - 0.5) < 5e-3 and abs(float(big.var()) - 1 / 12) < 2e-3
