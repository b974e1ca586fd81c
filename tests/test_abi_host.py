"""CPU-side checks: the C-ABI library builds/loads and exports every symbol include/tnerf.h declares,
the host-side mirror modules import, refuse CPU tensors loudly, and the pure-host helpers match the
reference vectors.  No kernel is launched here."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def built():
    import __graft_entry__ as g
    g.build()
    import _engine
    return _engine


def header_symbols():
    text = open(os.path.join(ROOT, "include", "tnerf.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(tnerf_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(built):
    lib = ctypes.CDLL(built.LIB_PATH)
    declared = header_symbols()
    assert len(declared) >= 25
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/tnerf.h but not exported"
    assert sorted(built.exported_symbols()) == declared, "ctypes table and header disagree"
    assert lib.tnerf_abi_version() == built.ABI_VERSION == 2


def test_sass_is_blackwell_native(built):
    import shutil
    import subprocess
    if not shutil.which("cuobjdump"):
        pytest.skip("cuobjdump not available")
    sass = subprocess.run(["cuobjdump", "-sass", built.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in sass
    for mnemonic in ("UTCHMMA", "LDTM", "STTM", "UBLKCP"):
        assert mnemonic in sass, f"{mnemonic} missing from SASS: tcgen05 / TMA path not compiled in"


def test_missing_library_fails_loudly(built, monkeypatch):
    monkeypatch.setattr(built, "_lib", None)
    monkeypatch.setattr(built, "LIB_PATH", "/nonexistent/libtnerf.so")
    with pytest.raises(RuntimeError, match="no fallback"):
        built.lib()


def test_cpu_tensors_are_rejected(built):
    from rays import get_rays
    from sampling import stratified_samples
    from encoding import PositionalEncoding
    from nerf import TinyNeRF
    from volume import volume_render
    with pytest.raises(RuntimeError, match="CUDA"):
        get_rays(4, 4, 10.0, torch.eye(4))
    with pytest.raises(RuntimeError, match="CUDA"):
        stratified_samples(2.0, 6.0, 8, torch.zeros(3, 3), torch.zeros(3, 3))
    with pytest.raises(RuntimeError, match="CUDA"):
        PositionalEncoding(4)(torch.zeros(3, 3))
    with pytest.raises(RuntimeError, match="CUDA"):
        TinyNeRF(27, 16, 2, 1)(torch.zeros(3, 27))
    with pytest.raises(RuntimeError, match="CUDA"):
        volume_render(torch.zeros(2, 4, 3), torch.zeros(2, 4, 1), torch.zeros(2, 4), torch.zeros(2, 3))


def test_module_surface_matches_reference(built):
    """names, constructor defaults, state_dict keys and shapes of SURVEY.md section 8b"""
    from encoding import PositionalEncoding
    from nerf import TinyNeRF
    enc = PositionalEncoding()
    assert (enc.num_freqs, enc.include_input, enc.out_dim) == (10, True, 63)
    assert PositionalEncoding(6).out_dim == 39 and PositionalEncoding(6, False).out_dim == 36
    m = TinyNeRF(63)
    assert (m.in_dim, m.hidden, m.depth, m.skip_at) == (63, 128, 4, 2)
    shapes = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    assert shapes == {"layers.0.weight": (128, 63), "layers.0.bias": (128,), "layers.1.weight": (128, 128),
                      "layers.1.bias": (128,), "layers.2.weight": (128, 191), "layers.2.bias": (128,),
                      "layers.3.weight": (128, 128), "layers.3.bias": (128,), "sigma.0.weight": (1, 128),
                      "sigma.0.bias": (1,), "rgb.0.weight": (3, 128), "rgb.0.bias": (3,)}
    assert sum(p.numel() for p in m.parameters()) == 66308


def test_default_init_consumes_rng_like_the_reference(built, golden):
    """torch.manual_seed(0); TinyNeRF(63,128,4,2) gives the reference's initial weights (make_golden.py)."""
    from nerf import TinyNeRF
    from encoding import PositionalEncoding
    torch.manual_seed(0)
    PositionalEncoding(10, True)
    m = TinyNeRF(63, 128, 4, 2)
    for k, v in m.state_dict().items():
        assert np.array_equal(v.numpy(), golden[f"mlp_repo_p_{k}"]), k


def test_spiral_poses_and_data_loader(built, golden, tmp_path):
    from camera import spiral_poses
    from data import load_tiny_nerf_npz
    out = spiral_poses(torch.from_numpy(golden["train_c2w"]), 7, 0.3)
    np.testing.assert_allclose(out.numpy(), golden["spiral"], rtol=1e-6, atol=1e-7)
    path = tmp_path / "d.npz"
    np.savez(path, images=np.zeros((2, 3, 3, 3), np.float32), poses=np.zeros((2, 4, 4), np.float64), focal=np.float64(138.8889))
    d = load_tiny_nerf_npz(str(path))
    assert d["poses"].dtype == np.float32 and d["focal"].dtype == np.float32 and d["images"].dtype == np.float32
    assert float(d["focal"]) == float(np.float32(138.8889))


def test_deferred_shape_logic(built):
    import _lazy
    spec = _lazy.SampleSpec(ro=None, o_stride=0, rd=torch.zeros(5, 3), n=5, S=8, near=2.0, far=6.0, near_t=None,
                            far_t=None, jitter=None, z_vals=torch.zeros(5, 8))
    d = _lazy.Deferred((5, 8, 3), torch.device("cpu"), _lazy.Node(spec), "pts")
    assert d.shape == (5, 8, 3) and d.shape[0] == 5 and d.dim() == 3 and d.size(1) == 8 and d.numel() == 120
    r = d.reshape(-1, 3)
    assert isinstance(r, _lazy.Deferred) and r.shape == (40, 3)
    assert isinstance(r.view(5, 8, 3), _lazy.Deferred) and isinstance(torch.reshape(r, (5, 8, 3)), _lazy.Deferred)
    assert _lazy._resolve_shape(d, (40, 3)) == [40, 3] and _lazy._resolve_shape(d, (-1,)) is None
    assert _lazy._resolve_shape(d, ((8, 5, 3),)) == [8, 5, 3] and _lazy._resolve_shape(d, (7, 3)) is None
