"""GPU parity tests of the fused hot path (tcgen05 kernels and the fp32 exact mode) against the CPU
oracle, with a SHARED jitter tensor.  Bars (BASELINE.json north_star): per-ray rgb/depth/acc within
2e-3 absolute, parameter gradients within 1e-2 relative (per tensor, rel-L2)."""
import math
import os

import pytest
import torch

from oracle import oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    return torch.device("cuda:0")


def rel_l2(a, b):
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def make_model(cfg, seed, dev, scale=1.0):
    from nerf import TinyNeRF
    ind, hid, dep, sk = cfg
    p = O.init_params(ind, hid, dep, sk, seed=seed)
    if scale != 1.0:   # "trained-like": larger weights -> non-trivial densities
        p = {k: (v * scale if k.endswith("weight") else v) for k, v in p.items()}
        p["sigma.0.bias"] = p["sigma.0.bias"] + 0.3
    m = TinyNeRF(ind, hid, dep, sk)
    m.load_state_dict(p)
    return m.to(dev), p


def random_rays(n, seed):
    g = torch.Generator().manual_seed(seed)
    pose = O.look_at_pose(2 * math.pi * float(torch.rand((), generator=g)), 0.3 + 0.4 * float(torch.rand((), generator=g)))
    ro, rd = O.get_rays(64, 64, 80.0, pose)
    idx = torch.randint(0, 64 * 64, (n,), generator=g)
    return ro[idx].contiguous(), rd[idx].contiguous()


def assert_render_parity(out, p, ro, rd, S, u, tol, pre_tol, white_bkgd=True, depth_tol=None, **kw):
    """Every ray of ``out`` = (comp, depth, acc) [depth / acc may be None] within ``tol`` ABSOLUTE of the oracle -- or of the
    oracle with the ray's LAST sample on the other side of the delta_last = 1e10 discontinuity (src/volume.py:20-23, SURVEY.md
    F8/H10), and then only if that sample's density pre-activation really is within rounding (``pre_tol``) of 0, so that an
    outlier cannot hide behind the discontinuity for any other reason.  Returns the number of rays on the far branch."""
    ref = O.render_rays(p, ro, rd, 2.0, 6.0, S, u, white_bkgd=white_bkgd, **kw)

    def worst(r):
        errs = [(out[0].detach().cpu() - r[0]).abs().amax(1)]
        if out[1] is not None:        # depth: same absolute bar unless a case states its own
            errs.append((out[1].detach().cpu() - r[1]).abs().reshape(-1) * (tol / (depth_tol or tol)))
        if out[2] is not None:
            errs.append((out[2].detach().cpu() - r[2]).abs().reshape(-1))
        return torch.stack(errs).amax(0)
    off = worst(ref) >= tol
    if off.any():
        e_flip = worst(O.render_rays_last_flipped(p, ro, rd, 2.0, 6.0, S, u, white_bkgd=white_bkgd, **kw))
        pre = O.last_sample_sigma_pre(p, ro, rd, 2.0, 6.0, S, u, **kw)
        assert bool((e_flip[off] < tol).all()), f"rays outside the bar on both branches: {worst(ref)[off & (e_flip >= tol)].tolist()}"
        assert bool((pre[off].abs() < pre_tol).all()), pre[off].tolist()
        assert off.float().mean() < 0.03, f"{int(off.sum())} of {off.numel()} rays on the far side of the discontinuity"
    return int(off.sum())


# ------------------------------------------------------------------------------------------ UMMA
@pytest.mark.parametrize("mode", [0, 1, 2, 3])
@pytest.mark.parametrize("NK", [(128, 64), (16, 32), (64, 128), (256, 16)])
def test_umma_selftest(dev, mode, NK):
    """The descriptor / layout conventions every fused kernel relies on."""
    import _engine as E
    N, K = NK
    g = torch.Generator().manual_seed(N * 7 + K + mode)
    a = torch.randn((K, 128) if mode == 3 else (128, K), generator=g)
    b = torch.randn((K, N) if mode == 2 else (N, K), generator=g)
    d = torch.full((128, N), float("nan"), device=dev)
    a_d, b_d = a.to(dev), b.to(dev)           # named: must outlive the asynchronous launch
    E.check(E.lib().tnerf_umma_selftest(E.ptr(a_d), E.ptr(b_d), N, K, mode, E.ptr(d), E.stream(dev)))
    A = a.half().float().t() if mode == 3 else a.half().float()
    B = b.half().float().t() if mode == 2 else b.half().float()
    ref = A @ B.t()
    assert (d.cpu() - ref).abs().max() < 1e-3 * max(1.0, ref.abs().max().item())


# ------------------------------------------------------------------------------------------ render
@pytest.mark.parametrize("prec", ["f32", "f16"])
@pytest.mark.parametrize("case", [
    dict(cfg=(63, 128, 4, 2), S=64, n=300, jitter=True, scale=1.0),
    dict(cfg=(63, 128, 4, 2), S=64, n=1031, jitter=False, scale=2.0),
    dict(cfg=(39, 128, 4, 2), S=32, n=257, jitter=True, scale=2.0),
    dict(cfg=(63, 128, 4, 2), S=192, n=70, jitter=False, scale=2.0),
    dict(cfg=(63, 128, 3, 1), S=128, n=65, jitter=True, scale=2.0),
    dict(cfg=(63, 128, 5, 0), S=96, n=50, jitter=True, scale=1.5),
    # n_samples not a multiple of 32: not covered by the tensor-core render kernel, the f16 request runs on the exact fp32 path
    dict(cfg=(60, 128, 2, 1), S=16, n=200, jitter=True, scale=2.0),
    # wide model of BASELINE config 4 (hidden 256): CTA-pair tcgen05 kernel on the f16 path
    dict(cfg=(63, 256, 4, 2), S=192, n=301, jitter=False, scale=1.5),
    dict(cfg=(63, 256, 4, 2), S=64, n=1000, jitter=True, scale=1.5),
    dict(cfg=(39, 256, 4, 2), S=32, n=3, jitter=True, scale=1.5),
    dict(cfg=(63, 256, 4, 2), S=96, n=1, jitter=True, scale=1.0),
])
def test_fused_render_vs_oracle(dev, prec, case):
    import engine
    from encoding import PositionalEncoding
    ind = case["cfg"][0]
    inc = (ind - 3) % 6 == 0
    L = (ind - 3) // 6 if inc else ind // 6
    enc = PositionalEncoding(L, inc).to(dev)
    model, p = make_model(case["cfg"], 5, dev, case["scale"])
    n, S = case["n"], case["S"]
    ro, rd = random_rays(n, 11)
    u = torch.rand(n, S, generator=torch.Generator().manual_seed(12)) if case["jitter"] else None
    with torch.no_grad():
        comp, depth, acc = engine.render_rays(model, enc, ro.to(dev), rd.to(dev), 2.0, 6.0, S,
                                              t_rand=None if u is None else u.to(dev), precision=prec)
    tol = 2e-5 if prec == "f32" else 2e-3          # BASELINE north_star: rgb / depth / acc within 2e-3 ABSOLUTE (depth too: no z_far factor)
    dtol = case.get("depth_tol_f16", tol) if prec == "f16" else tol
    assert_render_parity((comp, depth, acc), p, ro, rd, S, u, tol, 1e-5 if prec == "f32" else 4e-3, depth_tol=dtol,
                         num_freqs=L, include_input=inc, depth=case["cfg"][2], skip_at=case["cfg"][3])
    assert acc.min() >= 0 and comp.shape == (n, 3) and depth.shape == (n, 1)


@pytest.mark.parametrize("cfg,S,prec", [((63, 128, 4, 2), 64, "f16"), ((63, 256, 4, 2), 96, "f16"), ((63, 128, 4, 2), 24, "f16"),
                                        ((39, 128, 3, 1), 32, "f32")])
def test_pose_batched_frames_equal_per_pose_renders(dev, cfg, S, prec):
    """(f) N3: tnerf_render_frames (the make_gif.py frame loop as one call) returns exactly what one single-pose render per frame returns --
    one launch for the batch on the role-split kernels (n_samples % 32 == 0), a loop inside the call otherwise."""
    import _engine as E
    import engine
    from camera import spiral_poses
    from encoding import PositionalEncoding
    L = (cfg[0] - 3) // 6
    enc = PositionalEncoding(L, True).to(dev)
    model, p = make_model(cfg, 31, dev, 1.5)
    H, W, focal = 37, 53, 60.0
    path = spiral_poses(O.look_at_pose(0.2, 0.5).to(dev), n_frames=5, radius=0.4)
    l0 = E.launch_count()
    imgs, depth, acc = engine.render_frames(model, enc, H, W, focal, path, n_samples=S, precision=prec, return_aux=True)
    launches = E.launch_count() - l0
    assert imgs.shape == (5, H, W, 3) and depth.shape == (5, H, W, 1)
    if prec == "f16" and S % 32 == 0:
        assert launches <= 2, launches            # (pack +) ONE render launch for five frames
    import ctypes as C
    h = E.handle_for(model, dev)
    for i in range(5):          # the same frame as a single-pose call (rays generated in-kernel from the pose): bit-identical
        comp = torch.empty(H * W, 3, device=dev)
        pose_i = path[i].contiguous()
        rs = engine.ray_source(c2w=pose_i, H=H, W=W, focal=focal, first_ray=0)
        E.check(E.lib().tnerf_render_fwd(h.h, C.byref(rs), H * W, 2.0, 6.0, S, None, 1, engine._PREC[prec], E.ptr(comp), None, None, None, None,
                                         E.stream(dev)))
        assert torch.equal(comp.reshape(H, W, 3).clamp(0, 1), imgs[i]), i
    oro, ord_ = O.get_rays(H, W, focal, path[3].cpu())
    oc, _, _, _ = O.render_rays(p, oro, ord_, 2.0, 6.0, S, None, num_freqs=L, depth=cfg[2], skip_at=cfg[3])
    err = (imgs[3].reshape(-1, 3).cpu() - oc.clamp(0, 1)).abs().max(dim=1).values
    assert (err < (2e-5 if prec == "f32" else 2e-3)).float().mean() > 0.97


def test_config4_pair_kernel_against_reference_vectors(dev, golden_c4):
    """The CTA-pair kernel, through the C ABI with in-kernel ray generation, against pixels the UNMODIFIED reference produced for
    the config-4 model (tests/golden/make_golden_c4.py): 2e-3 on rgb / acc."""
    import ctypes as C
    import importlib.util
    import _engine as E
    import engine
    from nerf import TinyNeRF
    spec = importlib.util.spec_from_file_location("_make_golden_c4", os.path.join(os.path.dirname(__file__), "golden", "make_golden_c4.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    model = TinyNeRF(63, 256, 4, 2)
    model.load_state_dict(mod.synthetic_params())
    model = model.to(dev)
    h = E.handle_for(model, dev)
    h.set_encoding(10, True)
    h.ensure_packed(force=True)
    g = golden_c4
    pose = torch.from_numpy(g["c4_c2w"]).to(dev)
    pick = torch.from_numpy(g["c4_pick"]).to(dev)
    n = int(pick.numel())
    comp, depth, acc = torch.empty(n, 3, device=dev), torch.empty(n, 1, device=dev), torch.empty(n, 1, device=dev)
    rs = engine.ray_source(c2w=pose, H=800, W=800, focal=1111.11, pixel_index=pick)
    E.check(E.lib().tnerf_render_fwd(h.h, C.byref(rs), n, 2.0, 6.0, 192, None, 1, E.PREC_F16_TC, E.ptr(comp), E.ptr(depth), E.ptr(acc), None,
                                     None, E.stream(dev)))
    last_pre = torch.from_numpy(g["c4_sigma"]).reshape(n, 192)[:, -1]      # relu'd: rays whose last density is exactly 0 may sit on the discontinuity
    keep = last_pre > 4e-3
    assert keep.sum() >= n - 3
    assert (comp.cpu() - torch.from_numpy(g["c4_comp"]))[keep].abs().max() < 2e-3
    assert (acc.cpu() - torch.from_numpy(g["c4_acc"]))[keep].abs().max() < 2e-3
    assert (depth.cpu() - torch.from_numpy(g["c4_depth"]))[keep].abs().max() < 2e-3


def test_two_wide_models_alternate(dev):
    """The pair kernel keeps its epilogue biases in a constant-bank table owned by one (model, pack version) at a time: renders of two
    live hidden-256 models interleaved on two streams, and a re-packed model, must each see their own table."""
    import engine
    from encoding import PositionalEncoding
    enc = PositionalEncoding(10, True).to(dev)
    ma, pa = make_model((63, 256, 4, 2), 81, dev, 1.5)
    mb, pb = make_model((63, 256, 4, 2), 82, dev, 1.5)
    ro, rd = random_rays(500, 83)
    ro_d, rd_d = ro.to(dev), rd.to(dev)
    with torch.no_grad():
        ra = engine.render_rays(ma, enc, ro_d, rd_d, 2.0, 6.0, 64, precision="f16")[0].clone()
        rb = engine.render_rays(mb, enc, ro_d, rd_d, 2.0, 6.0, 64, precision="f16")[0].clone()
        assert (ra - rb).abs().max() > 1e-3                      # different models
        s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
        torch.cuda.synchronize()
        outs = []
        for k in range(6):
            with torch.cuda.stream(s1 if k % 2 == 0 else s2):
                m = ma if k % 2 == 0 else mb
                outs.append(engine.render_rays(m, enc, ro_d, rd_d, 2.0, 6.0, 64, precision="f16")[0])
        torch.cuda.synchronize()
        for k, o in enumerate(outs):
            assert torch.equal(o, ra if k % 2 == 0 else rb), k
        # in-place parameter change -> new pack version -> the table follows
        ma.layers[1].bias.add_(0.05)
        rc = engine.render_rays(ma, enc, ro_d, rd_d, 2.0, 6.0, 64, precision="f16")[0]
    pa2 = {k: v.detach().cpu() for k, v in ma.state_dict().items()}
    assert_render_parity((rc, None, None), pa2, ro, rd, 64, None, 2e-3, 4e-3)
    assert (rc - ra).abs().max() > 1e-4


def test_config4_frame_rows_vs_oracle(dev):
    """BASELINE config 4 shape (800x800 frame, 192 samples/ray, hidden 256) through the C ABI with rays generated in-kernel from the
    pose: two rows from the middle of the frame against the oracle, and the same rows rendered as part of a larger row block
    (what a rank of the row-sharded render does) must be bit-identical."""
    import ctypes as C
    import _engine as E
    import engine
    from encoding import PositionalEncoding
    PositionalEncoding(10, True).to(dev)
    model, p = make_model((63, 256, 4, 2), 77, dev, 1.5)
    H = W = 800
    focal, S = 1111.11, 192
    pose = O.look_at_pose(0.4, 0.55)
    h = E.handle_for(model, dev)
    h.set_encoding(10, True)
    h.ensure_packed(force=True)
    pose_d = pose.to(dev)

    def render(first, n):
        comp, depth, acc = torch.empty(n, 3, device=dev), torch.empty(n, 1, device=dev), torch.empty(n, 1, device=dev)
        rs = engine.ray_source(c2w=pose_d, H=H, W=W, focal=focal, first_ray=first)
        E.check(E.lib().tnerf_render_fwd(h.h, C.byref(rs), n, 2.0, 6.0, S, None, 1, E.PREC_F16_TC, E.ptr(comp), E.ptr(depth), E.ptr(acc),
                                         None, None, E.stream(dev)))
        return comp.cpu(), depth.cpu(), acc.cpu()

    first, n = 400 * W, 2 * W
    comp, depth, acc = render(first, n)
    ro, rd = O.get_rays(H, W, focal, pose)
    ro, rd = ro[first:first + n].contiguous(), rd[first:first + n].contiguous()
    assert_render_parity((comp, depth, acc), p, ro, rd, S, None, 2e-3, 4e-3)
    big, _, _ = render(390 * W, 20 * W + 37)             # a rank's row block (ragged end): same rays, different tiles / CTAs
    assert torch.equal(big[10 * W:12 * W], comp)


@pytest.mark.parametrize("prec", ["f32", "f16"])
def test_fused_render_black_background_and_broadcast_origin(dev, prec):
    import engine
    from encoding import PositionalEncoding
    from rays import get_rays
    enc = PositionalEncoding(10, True).to(dev)
    model, p = make_model((63, 128, 4, 2), 9, dev, 2.0)
    pose = O.look_at_pose(0.3, 0.6)
    ro, rd = get_rays(20, 30, 40.0, pose.to(dev))           # rays_o is a stride-0 view
    with torch.no_grad():
        comp, depth, acc = engine.render_rays(model, enc, ro, rd, 2.0, 6.0, 64, white_bkgd=False, precision=prec)
    oro, ord_ = O.get_rays(20, 30, 40.0, pose)
    assert_render_parity((comp, depth, acc), p, oro, ord_, 64, None, 2e-5 if prec == "f32" else 2e-3, 1e-5 if prec == "f32" else 4e-3, white_bkgd=False)


# ------------------------------------------------------------------------------------------ gradients
# fp16-operand gradients are checked at the batch sizes the reference trains with (train.py:23 n_rand=2048,
# BASELINE config 3: 4096): the dominant error is ReLU-mask flips of units whose pre-activation is within
# rounding of 0, an incoherent per-sample term that averages out against the coherent gradient sum.
@pytest.mark.parametrize("prec,case", [
    ("f32", dict(cfg=(63, 128, 4, 2), S=64, n=256)), ("f32", dict(cfg=(39, 128, 3, 1), S=32, n=100)),
    ("f32", dict(cfg=(63, 128, 4, 2), S=128, n=37)), ("f32", dict(cfg=(27, 64, 2, 1), S=24, n=90)),
    ("f16", dict(cfg=(63, 128, 4, 2), S=64, n=4096)), ("f16", dict(cfg=(63, 128, 4, 2), S=64, n=2049)),
    ("f16", dict(cfg=(63, 128, 4, 2), S=128, n=2048)), ("f16", dict(cfg=(39, 128, 4, 2), S=16, n=8192)),
    ("f16", dict(cfg=(63, 128, 4, 2), S=32, n=4096)), ("f16", dict(cfg=(39, 128, 3, 1), S=32, n=2048)),
])
def test_fused_train_grads_vs_oracle(dev, prec, case):
    import engine
    from encoding import PositionalEncoding
    L = (case["cfg"][0] - 3) // 6
    enc = PositionalEncoding(L, True).to(dev)
    model, p = make_model(case["cfg"], 21, dev, 2.0)
    n, S = case["n"], case["S"]
    ro, rd = random_rays(n, 22)
    g = torch.Generator().manual_seed(23)
    u, target = torch.rand(n, S, generator=g), torch.rand(n, 3, generator=g)
    comp, depth, acc = engine.render_rays(model, enc, ro.to(dev), rd.to(dev), 2.0, 6.0, S, t_rand=u.to(dev), precision=prec)
    loss = ((comp - target.to(dev)) ** 2).mean()
    loss.backward()
    l_ref, g_ref, _ = O.loss_and_grads(p, ro, rd, target, 2.0, 6.0, S, u, num_freqs=L, depth=case["cfg"][2], skip_at=case["cfg"][3])
    assert abs(loss.item() - l_ref.item()) < (1e-5 if prec == "f32" else 1e-3)
    tol = 1e-3 if prec == "f32" else 1e-2          # north-star bar: gradients within 1e-2 relative
    errs = {k: rel_l2(v.grad.cpu(), g_ref[k]) for k, v in model.named_parameters()}
    assert max(errs.values()) < tol, errs


@pytest.mark.parametrize("prec", ["f32", "f16"])
@pytest.mark.parametrize("white", [True, False])
def test_train_fwd_bwd_entry_point_vs_oracle(dev, prec, white):
    """tnerf_train_fwd_bwd (MSE inside the kernel, rays generated in-kernel from pose + pixel ids):
    loss, rendered colour and the flat gradient against the oracle's autograd."""
    import ctypes as C
    import _engine as E
    import engine
    from encoding import PositionalEncoding
    enc = PositionalEncoding(10, True).to(dev)
    model, p = make_model((63, 128, 4, 2), 51, dev, 2.0)
    H, W, focal, S = 40, 50, 60.0, 64
    n = 700 if prec == "f32" else 4096          # fp16 gradients are judged at the reference's batch size (see above)
    pose = O.look_at_pose(0.9, 0.45)
    g = torch.Generator().manual_seed(52)
    pix = torch.randint(0, H * W, (n,), generator=g)
    u, target = torch.rand(n, S, generator=g), torch.rand(n, 3, generator=g)
    h = E.handle_for(model, dev)
    h.set_encoding(10, True)
    h.ensure_packed(force=True)
    pose_d, pix_d, u_d, t_d = pose.to(dev), pix.to(dev), u.to(dev), target.to(dev)
    grads = torch.zeros(h.param_count, device=dev)
    loss = torch.zeros(1, device=dev)
    comp = torch.empty(n, 3, device=dev)
    rs = engine.ray_source(c2w=pose_d, H=H, W=W, focal=focal, pixel_index=pix_d)
    E.check(E.lib().tnerf_train_fwd_bwd(h.h, C.byref(rs), E.ptr(t_d), n, 2.0, 6.0, S, E.ptr(u_d), int(white), engine._PREC[prec],
                                        3.0 * n, E.ptr(comp), E.ptr(loss), E.ptr(grads), None, None, E.stream(dev)))
    oro, ord_ = O.get_rays(H, W, focal, pose)
    l_ref, g_ref, (oc, _, _) = O.loss_and_grads(p, oro[pix], ord_[pix], target, 2.0, 6.0, S, u, white_bkgd=white)
    assert_render_parity((comp, None, None), p, oro[pix], ord_[pix], S, u, 2e-5 if prec == "f32" else 2e-3, 1e-5 if prec == "f32" else 4e-3, white_bkgd=white)
    assert abs(loss.item() - l_ref.item()) < (1e-5 if prec == "f32" else 2e-3 * max(1.0, l_ref.item()))
    flat_ref = torch.cat([g_ref[k].reshape(-1) for k, _ in O.mlp_param_shapes(63, 128, 4, 2)])
    off = 0
    for k, shp in O.mlp_param_shapes(63, 128, 4, 2):
        cnt = g_ref[k].numel()
        err = rel_l2(grads[off:off + cnt].cpu(), flat_ref[off:off + cnt])
        assert err < (1e-3 if prec == "f32" else 1e-2), (k, err)
        off += cnt


def test_train_kernel_variants_agree_on_a_large_batch(dev):
    """The training kernel has two tile programs (rolled / unrolled per step) and two stream schedules (in phase = default; half a
    tile apart = option train_sync 0, the run-to-run reproducible mode: the two streams of a CTA then feed the shared weight-gradient
    accumulators in a fixed order).  On 16 384 rays x 64 samples: the two programs return the same gradient BITS under the
    reproducible schedule, the in-phase schedule agrees with it to fp32 summation-order noise, and the sum of two half batches
    matches the full batch."""
    import ctypes as C
    import _engine as E
    import engine
    from encoding import PositionalEncoding
    PositionalEncoding(10, True).to(dev)
    model, p = make_model((63, 128, 4, 2), 71, dev, 2.0)
    H, W, focal, S, n = 200, 200, 260.0, 64, 16384
    g = torch.Generator().manual_seed(72)
    pix = torch.randint(0, H * W, (n,), generator=g).to(dev)
    u, target = torch.rand(n, S, generator=g).to(dev), torch.rand(n, 3, generator=g).to(dev)
    pose = O.look_at_pose(1.1, 0.4).to(dev)
    h = E.handle_for(model, dev)
    h.set_encoding(10, True)
    h.ensure_packed(force=True)

    def run(lo, cnt, unroll_from, sync):
        h.set_option("unroll_from", unroll_from)
        h.set_option("train_sync", sync)
        try:
            grads = torch.zeros(h.param_count, device=dev)
            loss = torch.zeros(1, device=dev)
            rs = engine.ray_source(c2w=pose, H=H, W=W, focal=focal, pixel_index=pix[lo:lo + cnt].contiguous())
            E.check(E.lib().tnerf_train_fwd_bwd(h.h, C.byref(rs), E.ptr(target[lo:lo + cnt].contiguous()), cnt, 2.0, 6.0, S,
                                                E.ptr(u[lo:lo + cnt].contiguous()), 1, E.PREC_F16_TC, 3.0 * n, None, E.ptr(loss), E.ptr(grads),
                                                None, None, E.stream(dev)))
            torch.cuda.synchronize()
            return loss.cpu(), grads.cpu()
        finally:
            h.set_option("unroll_from", -1)
            h.set_option("train_sync", -1)

    l_u, g_u = run(0, n, 1, 0)               # unrolled tile program, reproducible schedule
    l_r, g_r = run(0, n, 1 << 30, 0)         # rolled
    l_r2, g_r2 = run(0, n, 1 << 30, 0)       # ... twice
    assert torch.equal(g_u, g_r) and torch.equal(g_r, g_r2), (g_u - g_r).abs().max().item()
    assert abs(l_u.item() - l_r.item()) < 1e-6          # the loss is summed with atomics across CTAs: last-bit differences between runs
    l_s, g_s = run(0, n, 1, 1)               # default: streams in phase
    assert rel_l2(g_s, g_r) < 1e-5 and abs(l_s.item() - l_r.item()) < 1e-6
    l_a, g_a = run(0, n // 2, 1 << 30, 0)
    l_b, g_b = run(n // 2, n // 2, 1 << 30, 0)
    assert abs((l_a + l_b).item() - l_u.item()) < 1e-5
    assert rel_l2(g_a + g_b, g_u) < 2e-3     # same rays, different tile -> CTA assignment and loss-scale rounding of the fp16 gradients


def test_trainer_steps_match_oracle_adam(dev):
    """fused optimisation steps (train kernel + slab reduce + Adam kernel + re-pack) vs oracle loss_and_grads +
    adam_step.  Before every step the engine is re-synchronised to the oracle's parameters and moments: Adam divides
    by sqrt(v), so entries whose gradient is summation noise (|g| ~ eps) move by up to lr in either direction on CPU
    and GPU alike and trajectories drift apart; the trajectory-level statement is the PSNR test."""
    import engine
    from encoding import PositionalEncoding
    enc = PositionalEncoding(10, True).to(dev)
    model, p = make_model((63, 128, 4, 2), 61, dev, 1.5)
    tr = engine.Trainer(model, enc, n_samples=32, precision="f32")
    names = [k for k, _ in O.mlp_param_shapes(63, 128, 4, 2)]
    m = {k: torch.zeros_like(v) for k, v in p.items()}
    v = {k: torch.zeros_like(x) for k, x in p.items()}
    H, W, focal, n, S = 30, 30, 45.0, 384, 32
    pose = O.look_at_pose(2.2, 0.35)
    oro, ord_ = O.get_rays(H, W, focal, pose)
    g = torch.Generator().manual_seed(62)
    for step in range(3):
        model.load_state_dict(p)
        tr.exp_avg.copy_(torch.cat([m[k].reshape(-1) for k in names]))
        tr.exp_avg_sq.copy_(torch.cat([v[k].reshape(-1) for k in names]))
        tr.steps = step
        tr.refresh()
        pix = torch.randint(0, H * W, (n,), generator=g)
        u, target = torch.rand(n, S, generator=g), torch.rand(n, 3, generator=g)
        loss = tr.step_pixels(pose.to(dev), H, W, focal, pix.to(dev), target.to(dev), u.to(dev))
        l_ref, g_ref, _ = O.loss_and_grads(p, oro[pix], ord_[pix], target, 2.0, 6.0, S, u)
        O.adam_step(p, g_ref, m, v, step + 1)
        assert abs(loss.item() - l_ref.item()) < 1e-5
        for k, prm in model.named_parameters():
            solid = g_ref[k].abs() > 1e-3 * g_ref[k].abs().max()
            diff = (prm.detach().cpu() - p[k]).abs()
            assert diff[solid].max() < 1e-5 and diff.max() <= 2.1 * 5e-4, (step, k, diff[solid].max().item(), diff.max().item())
    sd = tr.state_dict()
    assert sorted(sd["state"].keys()) == list(range(12)) and float(sd["state"][0]["step"]) == 3.0
    opt = torch.optim.Adam(model.parameters(), lr=5e-4)
    opt.load_state_dict(sd)                                    # interchangeable with torch.optim.Adam (checkpoints)


def test_gathering_optimizer_step_matches_the_scatter_path(dev, monkeypatch):
    """tnerf_train_fwd_bwd with grads = NULL leaves the (unscaled) gradient in the training kernel's sum vector and
    tnerf_optimizer_step (repack bit 1) gathers it: same parameters as the default path (scatter kernel + Adam) after a few steps --
    reproducible accumulation order is not available in that mode, so the comparison carries the last-bit tolerance of the sums --
    and an overflowed step leaves nothing behind in the vector."""
    import engine
    from encoding import PositionalEncoding
    enc = PositionalEncoding(10, True).to(dev)
    H, W, focal, n, S = 40, 40, 60.0, 2048, 64
    pose = O.look_at_pose(0.7, 0.45).to(dev)
    g = torch.Generator().manual_seed(77)
    batches = [(torch.randint(0, H * W, (n,), generator=g).to(dev), torch.rand(n, 3, generator=g).to(dev), torch.rand(n, S, generator=g).to(dev))
               for _ in range(4)]

    def run(gather):
        monkeypatch.setattr(engine, "_GATHER", gather)
        model, _ = make_model((63, 128, 4, 2), 78, dev, 1.5)
        tr = engine.Trainer(model, enc, n_samples=S)
        losses = []
        for k, (pix, tgt, jit) in enumerate(batches):
            if k == 2:      # a poisoned batch in between: skipped, and the sum vector must come out clean
                bad = tgt.clone(); bad[3, 0] = float("nan")
                tr.step_pixels(pose, H, W, focal, pix, bad, jit)
            losses.append(float(tr.step_pixels(pose, H, W, focal, pix, tgt, jit)))
        assert tr.applied_steps() == len(batches)
        return tr.flat.clone(), losses

    p_scatter, l_scatter = run(False)
    p_gather, l_gather = run(True)
    d = (p_scatter - p_gather).abs()
    assert d.max().item() <= 2e-4 and d.mean().item() <= 2e-6, (d.max().item(), d.mean().item())
    assert max(abs(a - b) for a, b in zip(l_scatter, l_gather)) < 1e-4


def test_fp32_trainer_does_not_leave_a_stale_operand_image(dev):
    """an fp32 Trainer writes the flat parameters from a raw kernel (no version-counter change); a later fp16 render
    must see the CURRENT weights, not the image packed at the first render."""
    import _engine as E
    import engine
    from encoding import PositionalEncoding
    enc = PositionalEncoding(10, True).to(dev)
    model, _ = make_model((63, 128, 4, 2), 81, dev, 1.5)
    H, W, focal, n, S = 24, 24, 36.0, 512, 32
    pose = O.look_at_pose(1.1, 0.4).to(dev)
    first = engine.render_frames(model, enc, H, W, focal, pose[None], n_samples=S)          # packs the image once
    tr = engine.Trainer(model, enc, n_samples=S, precision="f32", lr=5e-2)
    g = torch.Generator().manual_seed(82)
    for _ in range(3):
        tr.step_pixels(pose, H, W, focal, torch.randint(0, H * W, (n,), generator=g).to(dev), torch.rand(n, 3, generator=g).to(dev))
    after = engine.render_frames(model, enc, H, W, focal, pose[None], n_samples=S)
    E.handle_for(model, dev).ensure_packed(force=True)
    forced = engine.render_frames(model, enc, H, W, focal, pose[None], n_samples=S)
    assert torch.equal(after, forced)
    assert (after - first).abs().max() > 1e-3            # the weights did move


def test_tile_order_is_only_a_permutation_of_the_dealing(dev):
    """tnerf_set_tile_order / engine.Trainer.calibrate_tile_order (SM-speed-aware tile dealing): the calibration hands out a
    permutation of the CTAs, and ANY permutation -- the tiles of a step reach other CTAs, nothing else changes -- gives the same
    gradient and loss up to fp32 summation order, on a batch whose tiles do not divide evenly over the (CTA, stream) pairs."""
    import ctypes as C
    import _engine as E
    import engine
    from encoding import PositionalEncoding
    enc = PositionalEncoding(10, True).to(dev)
    model, _ = make_model((63, 128, 4, 2), 83, dev, 1.5)
    tr = engine.Trainer(model, enc, n_samples=64)
    summary = tr.calibrate_tile_order(force=True)
    sms = int(torch.cuda.get_device_properties(dev).multi_processor_count)
    assert sorted(tr.tile_order.cpu().tolist()) == list(range(sms))
    assert summary["ctas"] == sms and 0 < summary["loop_ns_min"] <= summary["loop_ns_median"] <= summary["loop_ns_max"]
    n, S = 4096 + 37, 64
    g = torch.Generator().manual_seed(84)
    pose = O.look_at_pose(0.7, 0.45).to(dev)
    pix = torch.randint(0, 100 * 100, (n,), generator=g).to(dev)
    tgt, jit = torch.rand(n, 3, generator=g).to(dev), torch.rand(n, S, generator=g).to(dev)
    rs = engine.ray_source(c2w=pose, H=100, W=100, focal=138.9, pixel_index=pix)

    def grad_and_loss(order):
        E.check(E.lib().tnerf_set_tile_order(tr.h.h, E.ptr(order), 0 if order is None else int(order.numel())), "tnerf_set_tile_order")
        out = torch.zeros(tr.P + 3, device=dev)
        loss_slot = out[tr.P:]
        E.check(E.lib().tnerf_train_fwd_bwd(tr.h.h, C.byref(rs), E.ptr(tgt), n, 2.0, 6.0, S, E.ptr(jit), 1, tr.prec, 3.0 * n, None,
                                            E.ptr(loss_slot), E.ptr(out), None, None, E.stream(dev)), "tnerf_train_fwd_bwd")
        torch.cuda.synchronize()
        return out[:tr.P + 1].clone()
    base = grad_and_loss(None)
    assert float(base[:-1].norm()) > 0 and float(base[-1]) > 0
    orders = [torch.arange(sms - 1, -1, -1, dtype=torch.int32, device=dev), torch.randperm(sms, generator=g).to(torch.int32).to(dev), tr.tile_order]
    for order in orders:
        got = grad_and_loss(order)
        assert float((got[:-1] - base[:-1]).norm() / base[:-1].norm()) < 1e-4
        assert abs(float(got[-1] - base[-1])) < 1e-4 * max(1.0, abs(float(base[-1])))


def test_optimizer_step_refreshes_operand_image_like_a_full_repack(dev):
    """tnerf_optimizer_step (Adam + clear gradient vector + in-place fp16 image refresh, one launch) against
    tnerf_adam_step followed by tnerf_pack_weights: same parameters, same moments, byte-identical image, cleared gradients."""
    import _engine as E
    import engine
    from encoding import PositionalEncoding
    for cfg, L in (((63, 128, 4, 2), 10), ((39, 128, 4, 2), 6)):
        enc = PositionalEncoding(L, True).to(dev)
        model, _ = make_model(cfg, 71, dev, 1.5)
        tr = engine.Trainer(model, enc, n_samples=64, grad_scaler=False)      # host-side step count: bit-comparable with tnerf_adam_step
        P = tr.P
        g = torch.Generator().manual_seed(72)
        grads = (torch.randn(P + 1, generator=g) * 1e-2).to(dev)
        ref_p, ref_m, ref_v = tr.flat.clone(), tr.exp_avg.clone(), tr.exp_avg_sq.clone()
        nbytes = E.lib().tnerf_packed_image_copy(tr.h.h, None, 0, None)
        assert nbytes > 0
        img_a, img_b = torch.zeros(nbytes, dtype=torch.uint8, device=dev), torch.zeros(nbytes, dtype=torch.uint8, device=dev)
        for step in (1, 2, 3):
            E.check(E.lib().tnerf_adam_step(E.ptr(ref_p), E.ptr(grads), E.ptr(ref_m), E.ptr(ref_v), P, step, 5e-4, 0.9, 0.999, 1e-8, 1.0, None,
                                            E.stream(dev)))
            tr.gbuf[:P + 1].copy_(grads)
            tr.steps = step - 1
            loss = tr._finish()
            assert torch.equal(tr.flat, ref_p) and torch.equal(tr.exp_avg, ref_m) and torch.equal(tr.exp_avg_sq, ref_v)
            assert float(tr.gbuf.abs().max()) == 0.0 and float(loss) == float(grads[P])
            assert E.lib().tnerf_packed_image_copy(tr.h.h, E.ptr(img_a), nbytes, E.stream(dev)) == nbytes
            tr.h.ensure_packed(force=True)                         # full re-pack from the fp32 parameters
            assert E.lib().tnerf_packed_image_copy(tr.h.h, E.ptr(img_b), nbytes, E.stream(dev)) == nbytes
            torch.cuda.synchronize()
            assert torch.equal(img_a, img_b), int((img_a != img_b).sum())


# ------------------------------------------------------------------------------------------ in-kernel jitter
@pytest.mark.parametrize("prec", ["f16", "f32"])
def test_in_kernel_jitter_equals_the_explicit_tensor_and_is_uniform(dev, prec):
    """src/sampling.py:24 draws the stratified jitter on the device; the training kernel draws it itself (counter-based Philox,
    tnerf_ray_source.jitter_seed/step) instead of reading a (N,S) tensor.  The numbers are a pure function of (seed, step, ray,
    sample): a step with the in-kernel draw equals the step fed with tnerf_jitter_fill's tensor -- which is also what the oracle
    gets -- and the draw is a sane uniform sample."""
    import engine
    from encoding import PositionalEncoding
    enc = PositionalEncoding(10, True).to(dev)
    H, W, focal, n, S = 40, 40, 60.0, 2048, 64
    pose = O.look_at_pose(1.3, 0.45)
    g = torch.Generator().manual_seed(97)
    pix, tgt = torch.randint(0, H * W, (n,), generator=g), torch.rand(n, 3, generator=g)

    def one_step(explicit):
        model, p = make_model((63, 128, 4, 2), 96, dev, 1.5)
        tr = engine.Trainer(model, enc, n_samples=S, precision=prec, jitter_seed=1234567)
        tr.h.set_option("train_sync", 0)                       # reproducible accumulation order: the comparison below is bit exact
        u = tr.jitter_tensor(n)
        loss = tr.step_pixels(pose.to(dev), H, W, focal, pix.to(dev), tgt.to(dev), u if explicit else None)
        tr.h.set_option("train_sync", -1)
        return tr.flat.clone(), float(loss), u.cpu(), p

    p_k, l_k, u, p0 = one_step(False)
    p_t, l_t, u2, _ = one_step(True)
    assert torch.equal(u, u2)
    if prec == "f16":      # same kernel, same numbers, fixed accumulation order: bit identical
        assert torch.equal(p_k, p_t) and abs(l_k - l_t) < 1e-6          # (the loss is summed with atomics across CTAs)
    else:                  # the fp32 path sums its split-K weight gradients with atomics: equal up to summation order
        assert float((p_k - p_t).abs().max()) < 1e-6 and abs(l_k - l_t) < 1e-6
    # statistics of the draw: uniform on [0, 1), no duplicates between steps / rays
    assert 0.0 <= float(u.min()) and float(u.max()) < 1.0 and abs(float(u.mean()) - 0.5) < 5e-3 and abs(float(u.var()) - 1 / 12) < 2e-3
    hist = torch.histc(u, bins=16, min=0.0, max=1.0) / u.numel()
    assert float((hist - 1 / 16).abs().max()) < 4e-3
    assert abs(float(torch.corrcoef(torch.stack([u[:, :-1].reshape(-1), u[:, 1:].reshape(-1)]))[0, 1])) < 1e-2
    # and the oracle, fed the same tensor, sees the same loss
    oro, ord_ = O.get_rays(H, W, focal, pose)
    l_ref, _, _ = O.loss_and_grads(p0, oro[pix], ord_[pix], tgt, 2.0, 6.0, S, u)
    assert abs(l_k - l_ref.item()) < (1e-5 if prec == "f32" else 2e-3 * max(1.0, l_ref.item()))


@pytest.mark.parametrize("case", [(1234567, 0, 257, 64), (0x9E3779B97F4A7C15, 3, 100, 128), (1, (1 << 32) + 17, 33, 24), ((1 << 63) + 5, 2 ** 40, 64, 1)])
def test_jitter_fill_is_philox_bit_for_bit(dev, case):
    """tnerf_jitter_fill (= the numbers the training kernel draws, previous test) against the oracle's Philox4x32-10 restatement,
    which is pinned to the published known-answer vectors (tests/test_oracle_golden.py::test_philox_known_answers): bit exact,
    including 64-bit seeds and step counts beyond 2^32."""
    import _engine as E
    seed, step, n, S = case
    out = torch.full((n, S), -1.0, device=dev)
    E.check(E.lib().tnerf_jitter_fill(seed, step, n, S, E.ptr(out), E.stream(dev)), "tnerf_jitter_fill")
    assert torch.equal(out.cpu(), O.jitter_uniform(seed, step, n, S))


# ------------------------------------------------------------------------------------------ GradScaler semantics
@pytest.mark.parametrize("prec", ["f16", "f32"])
def test_grad_scaler_skips_an_overflowed_step_and_recovers(dev, prec):
    """src/train.py:81,126-128 (GradScaler.scale / step / update) on the device: a step with a non-finite target must leave
    parameters, moments and Adam's step count untouched and halve the loss scale -- with no host synchronisation in between --
    and training must resume afterwards exactly as if the bad batch had never been seen."""
    import engine
    from encoding import PositionalEncoding
    enc = PositionalEncoding(10, True).to(dev)
    H, W, focal, n, S = 40, 40, 60.0, 1024, 64
    pose = O.look_at_pose(0.9, 0.4).to(dev)
    g = torch.Generator().manual_seed(91)
    batches = [(torch.randint(0, H * W, (n,), generator=g).to(dev), torch.rand(n, 3, generator=g).to(dev), torch.rand(n, S, generator=g).to(dev))
               for _ in range(4)]

    def run(with_bad_batch):
        model, _ = make_model((63, 128, 4, 2), 90, dev, 1.5)
        tr = engine.Trainer(model, enc, n_samples=S, precision=prec, growth_interval=1000)
        log = []
        for k, (pix, tgt, jit) in enumerate(batches):
            if with_bad_batch and k == 2:
                bad = tgt.clone(); bad[5, 1] = float("inf")
                before = (tr.flat.clone(), tr.exp_avg.clone(), tr.exp_avg_sq.clone(), float(tr.loss_scale), tr.applied_steps())
                loss = tr.step_pixels(pose, H, W, focal, pix, bad, jit)
                assert not math.isfinite(float(loss))
                assert torch.equal(tr.flat, before[0]) and torch.equal(tr.exp_avg, before[1]) and torch.equal(tr.exp_avg_sq, before[2])
                assert float(tr.loss_scale) == 0.5 * before[3] and tr.applied_steps() == before[4]
                assert float(tr.gbuf[:tr.P + 1].abs().max()) == 0.0          # the poisoned gradient vector was cleared all the same
            loss = tr.step_pixels(pose, H, W, focal, pix, tgt, jit)
            log.append(float(loss))
        assert tr.applied_steps() == len(batches) and all(math.isfinite(x) for x in log)
        sd = tr.state_dict()
        assert float(sd["state"][0]["step"]) == float(len(batches))
        return tr.flat.clone(), log, float(tr.loss_scale)

    p_clean, log_clean, scale_clean = run(False)
    p_bad, log_bad, scale_bad = run(True)
    assert scale_bad == 0.5 * scale_clean
    # same trajectory.  f16: the halved scale moves fp16 roundings of the backward operands.  f32: the split-K weight-gradient GEMMs
    # add their partials with atomics (order varies run to run: last-bit gradient differences).  Adam's m / sqrt(v) turns either into up
    # to a few per cent of one lr = 5e-4 step on the few parameters whose gradient is near zero (seen: 2e-6 ... 2.5e-5 on the f32 path
    # from one run to the next), so the statement is: no parameter further apart than 2e-4, and the bulk identical
    d = (p_clean - p_bad).abs()
    assert d.max().item() <= 2e-4 and d.mean().item() <= 2e-6, (d.max().item(), d.mean().item())
    assert max(abs(a - b) for a, b in zip(log_clean, log_bad)) < 1e-4


def test_grad_scaler_backs_off_from_a_too_large_scale_and_grows_again(dev):
    """a loss scale that pushes the head gradients out of the fp16-safe range is detected in the kernel (no infinities appear:
    the operands saturate): steps are skipped and the scale halves until the step fits; growth_interval clean steps double it."""
    import engine
    from encoding import PositionalEncoding
    enc = PositionalEncoding(10, True).to(dev)
    model, _ = make_model((63, 128, 4, 2), 93, dev, 1.5)
    H, W, focal, n, S = 40, 40, 60.0, 1024, 64
    pose = O.look_at_pose(0.2, 0.5).to(dev)
    g = torch.Generator().manual_seed(94)
    pix, tgt, jit = torch.randint(0, H * W, (n,), generator=g).to(dev), torch.rand(n, 3, generator=g).to(dev), torch.rand(n, S, generator=g).to(dev)
    tr = engine.Trainer(model, enc, n_samples=S, init_scale=2.0 ** 36, growth_interval=3)
    p0 = tr.flat.clone()
    scales = []
    for _ in range(40):
        tr.step_pixels(pose, H, W, focal, pix, tgt, jit)
        scales.append(float(tr.loss_scale))
    applied = tr.applied_steps()
    assert scales[0] == 2.0 ** 35 and 0 < applied < 40                # the first steps were skipped ...
    assert not torch.equal(tr.flat, p0) and bool(torch.isfinite(tr.flat).all())
    assert min(scales) < 2.0 ** 30 and any(b > a for a, b in zip(scales, scales[1:]))    # ... and the scale grows back after clean runs


# ------------------------------------------------------------------------------------------ deferred fusion
def test_reference_call_sequence_is_fused(dev):
    """src/train.py:114-121 verbatim call sequence -> one fused forward launch, same numbers."""
    import _engine as E
    from encoding import PositionalEncoding
    from sampling import stratified_samples
    from volume import volume_render
    enc = PositionalEncoding(10, True).to(dev)
    model, p = make_model((63, 128, 4, 2), 31, dev, 2.0)
    n, S = 512, 64
    ro, rd = random_rays(n, 32)
    u = torch.rand(n, S, generator=torch.Generator().manual_seed(33))
    target = torch.rand(n, 3, generator=torch.Generator().manual_seed(34))
    ro_d, rd_d = ro.to(dev), rd.to(dev)
    model.train()
    z_vals, pts = stratified_samples(2.0, 6.0, S, ro_d, rd_d, randomized=True, t_rand=u.to(dev))
    before = E.launch_count()
    with torch.amp.autocast("cuda", enabled=True):
        xenc = enc(pts.reshape(-1, 3))
        rgb, sigma = model(xenc)
        rgb = rgb.reshape(n, S, 3)
        sigma = sigma.reshape(n, S, 1)
        comp_rgb, _, _, w = volume_render(rgb, sigma, z_vals, rd_d)
        loss = torch.mean((comp_rgb - target.to(dev)) ** 2)
    launches_fwd = E.launch_count() - before
    assert launches_fwd <= 2, launches_fwd           # (weight pack +) one fused kernel
    scaler = torch.amp.GradScaler("cuda")
    scaler.scale(loss).backward()
    l_ref, g_ref, (oc, _, _) = O.loss_and_grads(p, ro, rd, target, 2.0, 6.0, S, u)
    keep = O.last_sample_sigma_pre(p, ro, rd, 2.0, 6.0, S, u).abs() > 4e-3
    assert_render_parity((comp_rgb.detach(), None, None), p, ro, rd, S, u, 2e-3, 4e-3)      # every ray, either branch of the discontinuity
    for k, v in model.named_parameters():
        assert rel_l2(v.grad.cpu() / scaler.get_scale(), g_ref[k]) < 1e-2, k
    assert w.shape == (n, S)
    ow = O.render_rays(p, ro, rd, 2.0, 6.0, S, u)[3]
    assert ((w + 0).cpu() - ow)[keep].abs().max() < 2e-3


def test_deferred_weights_refuse_to_describe_a_modified_model(dev):
    """The 4th output of a fused volume_render (per-sample weights) is recomputed on demand.  Asked for AFTER the parameters were
    modified in place it must fail loudly (as autograd does for a modified saved tensor), not describe the new network; asked for
    before, it is the oracle's weights."""
    from encoding import PositionalEncoding
    from sampling import stratified_samples
    from volume import volume_render
    enc = PositionalEncoding(10, True).to(dev)
    model, p = make_model((63, 128, 4, 2), 71, dev, 1.5)
    n, S = 64, 32
    ro, rd = random_rays(n, 72)
    ro_d, rd_d = ro.to(dev), rd.to(dev)

    def render():
        z_vals, pts = stratified_samples(2.0, 6.0, S, ro_d, rd_d, randomized=False)
        rgb, sigma = model(enc(pts.reshape(-1, 3)))
        return volume_render(rgb.reshape(n, S, 3), sigma.reshape(n, S, 1), z_vals, rd_d)
    with torch.no_grad():
        comp, _, _, w = render()
        comp = comp + 0
        w_now = w + 0                                      # read before the update: fine
        comp2, _, _, w2 = render()
        comp2 = comp2 + 0
        model.layers[0].bias.add_(0.01)                    # what optimizer.step() does
        with pytest.raises(RuntimeError, match="modified in place"):
            w2 + 0
    ow = O.render_rays(p, ro, rd, 2.0, 6.0, S, None)[3]
    clear = O.last_sample_sigma_pre(p, ro, rd, 2.0, 6.0, S, None).abs() > 4e-3
    assert w_now.shape == (n, S) and (w_now.cpu() - ow)[clear].abs().max() < 2e-3


def test_fused_backward_refuses_parameters_modified_since_the_forward(dev):
    """The fused backward recomputes the activations from the parameters; the reference's autograd graph raises when a parameter it
    needs was modified in place between forward and backward -- so does this one (instead of differentiating another network)."""
    import engine
    from encoding import PositionalEncoding
    enc = PositionalEncoding(10, True).to(dev)
    model, _ = make_model((63, 128, 4, 2), 73, dev, 1.5)
    ro, rd = random_rays(128, 74)
    comp, _, _ = engine.render_rays(model, enc, ro.to(dev), rd.to(dev), 2.0, 6.0, 64)
    with torch.no_grad():
        model.layers[1].weight.mul_(1.01)
    with pytest.raises(RuntimeError, match="modified by an inplace operation"):
        comp.sum().backward()
    comp, _, _ = engine.render_rays(model, enc, ro.to(dev), rd.to(dev), 2.0, 6.0, 64)     # a fresh forward differentiates fine
    comp.sum().backward()
    assert all(p.grad is not None and bool(torch.isfinite(p.grad).all()) for p in model.parameters())


def test_deferred_falls_back_when_chain_is_broken(dev):
    from encoding import PositionalEncoding
    from sampling import stratified_samples
    from volume import volume_render
    enc = PositionalEncoding(4, True).to(dev)
    model, p = make_model((27, 32, 3, 1), 41, dev, 2.0)     # hidden 32: no tensor-core path -> fp32 kernels
    n, S = 33, 24
    ro, rd = random_rays(n, 42)
    with torch.no_grad():
        z, pts = stratified_samples(2.0, 6.0, S, ro.to(dev), rd.to(dev), randomized=False)
        assert pts.shape == (n, S, 3) and pts.shape[0] == n
        rgb, sigma = model(enc(pts.reshape(-1, 3)))
        rgb = rgb.reshape(n, S, 3) * 1.0                   # arithmetic on a deferred tensor materialises it
        comp, depth, acc, w = volume_render(rgb, sigma.reshape(n, S, 1), z, rd.to(dev))
    oc, od, oa, ow = O.render_rays(p, ro, rd, 2.0, 6.0, S, None, num_freqs=4, depth=3, skip_at=1)
    assert (comp.cpu() - oc).abs().max() < 2e-5 and (w.cpu() - ow).abs().max() < 2e-5
    oz, op = O.stratified(2.0, 6.0, S, ro, rd, None)
    assert torch.equal(pts[3:5].cpu(), op[3:5])              # indexing a deferred tensor works too


def test_reference_call_sequence_with_the_wide_model(dev):
    """The same five-call sequence with TinyNeRF(hidden=256) (BASELINE config 4 model, 192 deterministic samples): the forward is
    one launch of the CTA-pair kernel; the per-sample weights (4th output) come from the fp32 path on demand."""
    import _engine as E
    from encoding import PositionalEncoding
    from sampling import stratified_samples
    from volume import volume_render
    enc = PositionalEncoding(10, True).to(dev)
    model, p = make_model((63, 256, 4, 2), 61, dev, 1.5)
    n, S = 150, 192
    ro, rd = random_rays(n, 62)
    ro_d, rd_d = ro.to(dev), rd.to(dev)
    model.eval()
    E.handle_for(model, dev).ensure_packed(force=True) if E.handle_for(model, dev).fused_ok else None
    with torch.no_grad():
        z_vals, pts = stratified_samples(2.0, 6.0, S, ro_d, rd_d, randomized=False)
        before = E.launch_count()
        rgb, sigma = model(enc(pts.reshape(-1, 3)))
        comp, depth, acc, w = volume_render(rgb.reshape(n, S, 3), sigma.reshape(n, S, 1), z_vals, rd_d)
        comp = comp + 0
        launches = E.launch_count() - before
        assert launches <= 2, launches               # (weight pack +) the pair kernel
        w = w + 0                                    # materialises the weights: fp32 path
    # every ray within the bar of the oracle on one side of the delta_last discontinuity or the other (no ray set aside unexamined)
    assert_render_parity((comp, depth, acc), p, ro, rd, S, None, 2e-3, 4e-3)
    oc, od, oa, ow = O.render_rays(p, ro, rd, 2.0, 6.0, S, None)
    clear = O.last_sample_sigma_pre(p, ro, rd, 2.0, 6.0, S, None).abs() > 4e-3      # the fp32 weights: compared where both sides agree on the branch
    assert w.shape == (n, S) and (w.cpu() - ow)[clear].abs().max() < 2e-4
