"""The oracle against the UNMODIFIED reference modules themselves, live, on inputs the golden fixtures do not contain (other seeds,
shapes, poses, sample counts).  The reference modules are the copies staged git-ignored under oracle/_ref/src by
tools/stage_reference.sh (run by __graft_entry__.build() where /root/reference exists; they travel to the GPU box with the
snapshot).  CPU only; skipped when the staged copies are absent -- tests/test_oracle_golden.py then carries the pinning alone."""
import importlib.util
import math
import os
import types

import numpy as np
import pytest
import torch

from oracle import oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_DIR = os.path.join(ROOT, "oracle", "_ref", "src")
NAMES = ["rays", "sampling", "encoding", "nerf", "volume", "utils"]

pytestmark = pytest.mark.skipif(not all(os.path.exists(os.path.join(REF_DIR, n + ".py")) for n in NAMES),
                                reason="reference modules not staged under oracle/_ref/src (run __graft_entry__.build() next to /root/reference)")


@pytest.fixture(scope="module")
def R():
    mods = {}
    for n in NAMES:                      # private names: the product package has modules of the same names on sys.path
        spec = importlib.util.spec_from_file_location("live_reference_" + n, os.path.join(REF_DIR, n + ".py"))
        m = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(m)
        mods[n] = m
    return types.SimpleNamespace(**mods)


def close(a, b, rtol=0.0, atol=0.0):
    np.testing.assert_allclose(a.detach().numpy(), b.detach().numpy(), rtol=rtol, atol=atol)


def some_pose(seed):
    g = torch.Generator().manual_seed(seed)
    th, ph = 2 * math.pi * float(torch.rand((), generator=g)), 0.2 + 0.9 * float(torch.rand((), generator=g))
    return O.look_at_pose(th, ph, 3.0 + 2.0 * float(torch.rand((), generator=g)))


@pytest.mark.parametrize("H,W,focal,seed", [(7, 5, 9.5, 1), (33, 64, 70.25, 2), (100, 100, 138.88888549804688, 3), (1, 1, 1.0, 4)])
def test_get_rays(R, H, W, focal, seed):
    pose = some_pose(seed)
    ro, rd = O.get_rays(H, W, focal, pose)
    rro, rrd = R.rays.get_rays(H, W, focal, pose)
    assert ro.shape == rro.shape == (H * W, 3)
    close(ro, rro)
    close(rd, rrd, rtol=1e-6, atol=1e-7)


@pytest.mark.parametrize("S", [1, 2, 3, 16, 64, 65, 128, 192, 257])
def test_stratified_deterministic_is_bit_exact(R, S):
    g = torch.Generator().manual_seed(10 + S)
    ro, rd = torch.randn(37, 3, generator=g), torch.randn(37, 3, generator=g)
    z, pts = O.stratified(2.0, 6.0, S, ro, rd, None)
    rz, rpts = R.sampling.stratified_samples(2.0, 6.0, S, ro, rd, randomized=False)
    assert torch.equal(z, rz.contiguous()) and torch.equal(pts, rpts)


@pytest.mark.parametrize("S,near,far", [(2, 2.0, 6.0), (8, 0.5, 3.25), (64, 2.0, 6.0), (96, 1.0, 7.5)])
def test_stratified_jittered_is_bit_exact_given_the_same_uniforms(R, S, near, far):
    """the reference draws torch.rand_like(z_vals) (src/sampling.py:24): the same generator state gives the oracle the same uniforms"""
    g = torch.Generator().manual_seed(20 + S)
    ro, rd = torch.randn(29, 3, generator=g), torch.randn(29, 3, generator=g)
    torch.manual_seed(1000 + S)
    rz, rpts = R.sampling.stratified_samples(near, far, S, ro, rd, randomized=True)
    torch.manual_seed(1000 + S)
    u = torch.rand(29, S)                # rand_like of a (29, S) fp32 CPU tensor consumes the global generator in the same way
    z, pts = O.stratified(near, far, S, ro, rd, u)
    assert torch.equal(z, rz) and torch.equal(pts, rpts)


@pytest.mark.parametrize("L,inc", [(1, True), (4, False), (6, True), (10, True), (12, False)])
def test_posenc_is_bit_exact(R, L, inc):
    x = torch.randn(211, 3, generator=torch.Generator().manual_seed(30 + L)) * 3.0
    enc = R.encoding.PositionalEncoding(L, inc)
    assert enc.out_dim == O.posenc_dim(L, inc)
    assert torch.equal(O.posenc(x, L, inc), enc(x))


@pytest.mark.parametrize("cfg", [(63, 128, 4, 2), (63, 256, 4, 2), (39, 64, 3, 1), (27, 32, 5, 3), (60, 128, 2, 1), (63, 128, 4, 0)])
def test_mlp_forward_with_the_reference_modules_own_initialisation(R, cfg):
    ind, hid, dep, sk = cfg
    torch.manual_seed(40 + hid + dep)
    ref = R.nerf.TinyNeRF(ind, hid, dep, sk)
    p = {k: v.detach().clone() for k, v in ref.state_dict().items()}
    assert [(k, tuple(v.shape)) for k, v in p.items()] == [(k, tuple(s)) for k, s in O.mlp_param_shapes(ind, hid, dep, sk)]
    x = torch.randn(301, ind, generator=torch.Generator().manual_seed(41))
    c, s = O.mlp_forward(p, x, dep, sk)
    with torch.no_grad():
        rc, rs = ref(x)
    close(c, rc, rtol=2e-6, atol=2e-7)
    close(s, rs, rtol=2e-6, atol=2e-6)


@pytest.mark.parametrize("S,white", [(2, True), (5, False), (64, True), (192, True)])
def test_volume_render_forward_and_backward(R, S, white):
    g = torch.Generator().manual_seed(50 + S)
    n = 23
    rgb = torch.rand(n, S, 3, generator=g)
    sigma = (torch.randn(n, S, 1, generator=g) * 2.0).clamp_min(0)          # a good share of exact zeros, like a ReLU output
    sigma[1] = 0.0                                                           # empty ray
    sigma[2] = 50.0                                                          # opaque ray
    z = torch.sort(2.0 + 4.0 * torch.rand(n, S, generator=g), dim=1).values
    rd = torch.randn(n, 3, generator=g)
    c, d, a, w = O.composite(rgb, sigma, z, rd, white)
    rgb_r, sig_r = rgb.clone().requires_grad_(True), sigma.clone().requires_grad_(True)
    rc, rdep, ra, rw = R.volume.volume_render(rgb_r, sig_r, z, rd, white_bkgd=white)
    close(c, rc, rtol=1e-6, atol=1e-7)
    close(d, rdep, rtol=1e-6, atol=1e-7)
    close(a, ra, rtol=1e-6, atol=1e-7)
    close(w, rw, rtol=1e-6, atol=1e-12)
    gC, gD, gA = torch.randn(n, 3, generator=g), torch.randn(n, 1, generator=g), torch.randn(n, 1, generator=g)
    (rc * gC).sum().add((rdep * gD).sum()).add((ra * gA).sum()).backward()
    d_rgb, d_sig = O.composite_backward(rgb.double(), sigma.double(), z.double(), rd.double(), gC.double(), gD.double(), gA.double(), None, white)
    close(d_rgb.float(), rgb_r.grad, rtol=1e-5, atol=1e-7)
    # fp32 autograd of cumprod divides by q: compare against the size of the gradient, not element by element
    ref = sig_r.grad.reshape(n, S).double()
    assert float((d_sig.reshape(n, S) - ref).abs().max()) <= 2e-4 * float(ref.abs().max()) + 1e-9


def test_single_sample_rays_are_a_documented_deviation(R):
    """n_samples = 1 is degenerate in the reference: src/volume.py:19-21 builds delta_inf from `deltas[..., :1]` of an EMPTY tensor, so
    there is no delta at all, every per-sample tensor broadcasts to (N, 0) and the ray shows the background only (weights (N, 0)).
    The oracle and the kernels composite the single sample as a LAST sample (delta = 1e10), the evident intent.  No BASELINE config
    uses one sample per ray; DESIGN.md section 6 lists the deviation.  This test pins both behaviours so that neither drifts."""
    n = 6
    g = torch.Generator().manual_seed(70)
    rgb, sigma = torch.rand(n, 1, 3, generator=g), torch.full((n, 1, 1), 5.0)
    z, rd = torch.full((n, 1), 2.0), torch.randn(n, 3, generator=g)
    rc, rdep, ra, rw = R.volume.volume_render(rgb, sigma, z, rd, white_bkgd=True)
    assert tuple(rw.shape) == (n, 0) and torch.equal(rc, torch.ones(n, 3)) and float(ra.abs().max()) == 0.0 and float(rdep.abs().max()) == 0.0
    c, d, a, w = O.composite(rgb, sigma, z, rd, True)
    assert tuple(w.shape) == (n, 1) and torch.allclose(a, torch.ones(n, 1)) and torch.allclose(c, rgb[:, 0, :], atol=1e-6)


def test_psnr(R):
    m = torch.tensor([1.0, 0.1, 3.3e-4, 1e-9])
    close(O.mse2psnr(m), R.utils.mse2psnr(m), rtol=1e-6)


def test_one_training_step_loss_and_gradients(R):
    """src/train.py:108-123,126 with the reference's modules and autograd against O.loss_and_grads on the same rays, targets and
    uniforms (fp32, no autocast -- the CPU path of the reference)."""
    torch.manual_seed(60)
    enc = R.encoding.PositionalEncoding(10, True)
    ref = R.nerf.TinyNeRF(enc.out_dim, 128, 4, 2)
    p = {k: v.detach().clone() for k, v in ref.state_dict().items()}
    n, S = 96, 32
    ro_all, rd_all = R.rays.get_rays(40, 40, 55.0, some_pose(61))
    g = torch.Generator().manual_seed(62)
    idx = torch.randint(0, 1600, (n,), generator=g)
    ro, rd, target = ro_all[idx], rd_all[idx], torch.rand(n, 3, generator=g)
    torch.manual_seed(63)
    z_vals, pts = R.sampling.stratified_samples(2.0, 6.0, S, ro, rd, randomized=True)
    torch.manual_seed(63)
    u = torch.rand(n, S)
    rgb, sigma = ref(enc(pts.reshape(-1, 3)))
    comp, depth, acc, _ = R.volume.volume_render(rgb.reshape(n, S, 3), sigma.reshape(n, S, 1), z_vals, rd)
    loss = torch.mean((comp - target) ** 2)
    loss.backward()
    l, grads, (oc, od, oa) = O.loss_and_grads(p, ro, rd, target, 2.0, 6.0, S, u)
    close(l, loss, rtol=1e-5)
    close(oc, comp, rtol=1e-5, atol=1e-6)
    close(od, depth, rtol=1e-5, atol=1e-6)
    close(oa, acc, rtol=1e-5, atol=1e-6)
    for k, v in ref.named_parameters():
        gr = v.grad
        assert float((grads[k] - gr).abs().max()) <= 1e-4 * float(gr.abs().max()) + 1e-9, k


# ------------------------------------------------------------------------------------------ the drop-in's call surface
def test_mirror_modules_have_the_reference_call_surface(R):
    """the product's flat modules (tiny-nerf-pytorch_b200/*.py) expose the reference's public functions and
    classes with the same parameter names, order and defaults -- checked against the staged reference modules by introspection
    (nothing is launched; a CPU tensor would be refused)."""
    import inspect

    import camera as m_camera
    import data as m_data
    import encoding as m_encoding
    import nerf as m_nerf
    import rays as m_rays
    import sampling as m_sampling
    import utils as m_utils
    import volume as m_volume
    extra = {}
    for n in ("camera", "data"):
        spec = importlib.util.spec_from_file_location("live_reference_" + n, os.path.join(REF_DIR, n + ".py"))
        extra[n] = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(extra[n])

    def params(fn):
        return [(p.name, p.default, p.kind) for p in inspect.signature(fn).parameters.values()]
    pairs = [(m_rays, R.rays), (m_sampling, R.sampling), (m_encoding, R.encoding), (m_nerf, R.nerf), (m_volume, R.volume), (m_utils, R.utils),
             (m_camera, extra["camera"]), (m_data, extra["data"])]
    checked = []
    for mine, ref in pairs:
        for name, obj in vars(ref).items():
            if name.startswith("_") or getattr(obj, "__module__", None) != ref.__name__:
                continue                                    # imported names (torch, np, ...) are not the module's surface
            assert hasattr(mine, name), f"{ref.__name__}.{name} has no counterpart"
            got = getattr(mine, name)
            if inspect.isclass(obj):
                assert params(got.__init__) == params(obj.__init__), name
                assert params(got.forward)[:len(params(obj.forward))] == params(obj.forward), name
            elif inspect.isfunction(obj):
                mine_p, ref_p = params(got), params(obj)
                # the mirror may ADD trailing keyword parameters (e.g. an explicit jitter tensor for parity runs), never change the reference's
                assert mine_p[:len(ref_p)] == ref_p, (name, mine_p, ref_p)
                assert all(p[1] is not inspect.Parameter.empty for p in mine_p[len(ref_p):]), name
            checked.append(name)
    assert {"get_rays", "stratified_samples", "PositionalEncoding", "TinyNeRF", "volume_render"} <= set(checked), checked


def test_train_script_keeps_the_reference_cli(R):
    """tiny-nerf-pytorch_b200/train.py keeps the reference's Config fields (names, types, defaults -> the same tyro command line,
    src/train.py:21-34), in the same order; it may add fields behind them.  Compared by parsing both files (neither is imported)."""
    import ast
    ref_path = os.path.join(ROOT, "baseline", "_ref", "src", "train.py")
    if not os.path.exists(ref_path):
        pytest.skip("reference scripts not staged under baseline/_ref/src")

    def fields(path):
        for node in ast.walk(ast.parse(open(path).read())):
            if isinstance(node, ast.ClassDef) and node.name == "Config":
                return [(s.target.id, ast.unparse(s.annotation), ast.unparse(s.value) if s.value else None)
                        for s in node.body if isinstance(s, ast.AnnAssign)]
        raise AssertionError(f"no Config in {path}")
    ref, mine = fields(ref_path), fields(os.path.join(ROOT, "tiny-nerf-pytorch_b200", "train.py"))
    assert len(ref) >= 10 and mine[:len(ref)] == ref
    assert all(default is not None for _, _, default in mine[len(ref):])


# ------------------------------------------------------------------------------------------ BASELINE config 1 (tiny_nerf_min.py)
@pytest.fixture(scope="module")
def tiny_min(tmp_path_factory):
    """src/tiny_nerf_min.py is a self-contained duplicate of the modules whose import has side effects (it prints,
    creates outputs/ and checkpoints/ in the working directory and builds its global model): imported from the staged copy inside a
    scratch directory, with the repo's imageio shim on the path."""
    import sys
    path = os.path.join(REF_DIR, "tiny_nerf_min.py")
    if not os.path.exists(path):
        pytest.skip("tiny_nerf_min.py not staged")
    shim = os.path.join(ROOT, "tiny-nerf-pytorch_b200", "_shims")
    cwd = os.getcwd()
    scratch = tmp_path_factory.mktemp("tiny_min")
    os.makedirs(scratch / "data")
    np.savez(scratch / "data" / "tiny_nerf_data.npz", **O.synthetic_scene(n_views=3, H=8, W=8, focal=11.0))     # the script loads it at import
    os.chdir(scratch)
    added = shim not in sys.path
    if added:
        sys.path.append(shim)                  # behind everything else: a real imageio wins
    try:
        spec = importlib.util.spec_from_file_location("live_reference_tiny_nerf_min", path)
        m = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(m)
    finally:
        os.chdir(cwd)
        if added:
            sys.path.remove(shim)
    return m


def test_config1_script_computes_what_the_modules_compute(tiny_min):
    """BASELINE config 1 is tiny_nerf_min.py's own train step (2048 rays x 64 samples): its private copies of the five stages against
    the oracle -- the same math as the modules', so `bench.py --workload c1` measures it with the same kernels."""
    M = tiny_min
    assert (M.N_RAND, M.N_SAMPLES, M.NEAR, M.FAR, M.LR) == (2048, 64, 2.0, 6.0, 5e-4)
    assert (M.enc.num_freqs, M.enc.include_input, M.enc.out_dim) == (10, True, 63)        # L = 10 as coded
    pose = some_pose(80)
    ro, rd = O.get_rays(21, 17, 30.5, pose)
    mro, mrd = M.get_rays(21, 17, 30.5, pose, torch.device("cpu"))
    close(ro, mro)
    close(rd, mrd, rtol=1e-6, atol=1e-7)
    g = torch.Generator().manual_seed(81)
    idx = torch.randint(0, 21 * 17, (64,), generator=g)
    ro, rd = ro[idx], rd[idx]
    torch.manual_seed(82)
    mz, mpts = M.stratified_samples(2.0, 6.0, 64, ro, rd, randomized=True)
    torch.manual_seed(82)
    u = torch.rand(64, 64)
    z, pts = O.stratified(2.0, 6.0, 64, ro, rd, u)
    assert torch.equal(z, mz) and torch.equal(pts, mpts)
    enc = M.PositionalEncoding(10, True)
    feat = O.posenc(pts.reshape(-1, 3), 10, True)
    assert torch.equal(feat, enc(pts.reshape(-1, 3)))
    torch.manual_seed(83)
    model = M.TinyNeRF(63, 128, 4, 2)
    p = {k: v.detach().clone() for k, v in model.state_dict().items()}
    assert [(k, tuple(v.shape)) for k, v in p.items()] == [(k, tuple(s)) for k, s in O.mlp_param_shapes(63, 128, 4, 2)]
    target = torch.rand(64, 3, generator=g)
    rgb, sigma = model(enc(mpts.reshape(-1, 3)))
    comp, depth, acc, w = M.volume_render(rgb.reshape(64, 64, 3), sigma.reshape(64, 64, 1), mz, rd)
    loss = torch.mean((comp - target) ** 2)
    loss.backward()
    l, grads, (oc, od, oa) = O.loss_and_grads(p, ro, rd, target, 2.0, 6.0, 64, u)
    close(l, loss, rtol=1e-5)
    close(oc, comp, rtol=1e-5, atol=1e-6)
    close(od, depth, rtol=1e-5, atol=1e-6)
    close(oa, acc, rtol=1e-5, atol=1e-6)
    for k, v in model.named_parameters():
        assert float((grads[k] - v.grad).abs().max()) <= 1e-4 * float(v.grad.abs().max()) + 1e-9, k
    close(O.mse2psnr(loss.detach()), M.mse2psnr(loss.detach()), rtol=1e-6)


# ------------------------------------------------------------------------------------------ edge cases the fixtures do not hold
def test_edge_cases_agree_bit_for_bit(R):
    """per-ray near / far tensors with jitter, a scaled (non-orthonormal) pose, L = 0, never-taken and early skips, unsorted depths, a
    zero-length direction, both backgrounds, enormous and infinite densities: the oracle equals the reference exactly on all of them"""
    g = torch.Generator().manual_seed(90)
    ro, rd = torch.randn(9, 3, generator=g), torch.randn(9, 3, generator=g)
    near = torch.rand(9, 1, generator=g) + 1.0
    far = near + 1.0 + 4.0 * torch.rand(9, 1, generator=g)
    torch.manual_seed(91)
    rz, rpts = R.sampling.stratified_samples(near, far, 12, ro, rd, True)
    torch.manual_seed(91)
    z, pts = O.stratified(near, far, 12, ro, rd, torch.rand(9, 12))
    assert torch.equal(z, rz) and torch.equal(pts, rpts)
    pose = some_pose(92)
    pose[:3, :3] *= 2.5
    (o1, d1), (o2, d2) = O.get_rays(6, 5, 7.0, pose), R.rays.get_rays(6, 5, 7.0, pose)
    assert torch.equal(o1, o2) and torch.allclose(d1, d2, rtol=1e-6, atol=1e-7)
    x = torch.randn(4, 3, generator=g)
    assert torch.equal(O.posenc(x, 0, True), R.encoding.PositionalEncoding(0, True)(x))
    for cfg in [(27, 16, 3, 1), (27, 16, 3, 5)]:                 # skip after the first layer / never
        torch.manual_seed(93)
        ref = R.nerf.TinyNeRF(*cfg)
        p = {k: v.detach().clone() for k, v in ref.state_dict().items()}
        xin = torch.randn(7, 27, generator=g)
        with torch.no_grad():
            rc, rs = ref(xin)
        c, s = O.mlp_forward(p, xin, cfg[2], cfg[3])
        close(c, rc, rtol=2e-6, atol=2e-7)
        close(s, rs, rtol=2e-6, atol=2e-6)
    with pytest.raises(RuntimeError):                            # skip_at == depth: the heads get hidden + in_dim features
        R.nerf.TinyNeRF(27, 16, 1, 1)(torch.zeros(2, 27))
    n, S = 5, 6
    rgb, sig = torch.rand(n, S, 3, generator=g), 3.0 * torch.rand(n, S, 1, generator=g)
    zz = 2.0 + 4.0 * torch.rand(n, S, generator=g)               # NOT sorted: negative deltas
    dirs = torch.randn(n, 3, generator=g)
    dirs[0] = 0.0                                                # zero-length direction: every delta is 0
    for white in (True, False):
        for a, b in zip(O.composite(rgb, sig, zz, dirs, white), R.volume.volume_render(rgb, sig, zz, dirs, white)):
            assert torch.equal(a, b)
    sig[1], sig[2] = 1e30, float("inf")
    zs = torch.sort(zz, dim=1).values
    for a, b in zip(O.composite(rgb, sig, zs, dirs, True), R.volume.volume_render(rgb, sig, zs, dirs, True)):
        assert torch.allclose(a, b, rtol=0, atol=0, equal_nan=True)
    close(O.mse2psnr(torch.tensor(-1.0)), R.utils.mse2psnr(torch.tensor(-1.0)))       # clamped at 1e-10: 100 dB
