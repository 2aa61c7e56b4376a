"""CPU checks of the drop-in boundary: libnerfpp_b200.so builds, loads, and exports exactly what include/nerfpp_b200.h
declares (no compute calls — there is no GPU here), and argument validation fails loudly."""
import ctypes
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


def declared_symbols():
    text = (ROOT / "include" / "nerfpp_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(nrf_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported_and_bound():
    from nerfpp_b200 import cabi
    lib = cabi.lib()
    names = declared_symbols()
    assert len(names) >= 24
    for n in names:
        assert hasattr(lib, n), f"{n} declared in the header but not exported"
    assert sorted(cabi.SIGNATURES) == names, "ctypes table and header disagree"
    assert lib.nrf_abi_version() == 1


def test_no_torch_in_the_abi_library():
    """The boundary is plain C: the shared library must not link torch / c10 / python."""
    import subprocess
    from nerfpp_b200 import build
    out = subprocess.run(["ldd", str(build.build())], capture_output=True, text=True).stdout
    assert "libtorch" not in out and "libc10" not in out and "libpython" not in out
    assert "libcudart" in out


def test_argument_validation_without_gpu():
    from nerfpp_b200 import cabi, ops
    lib = cabi.lib()
    assert lib.nrf_sh_encode_fwd(None, 3, 8, 9, None, None) == -1          # degree out of range
    assert b"degree" in lib.nrf_last_error()
    assert lib.nrf_sh_encode_fwd(None, 3, 0, 4, None, None) == 0           # empty input is a no-op
    shape = ops.mlp_shape(hidden=128)
    assert lib.nrf_mlp_small_param_count(ctypes.byref(shape)) == -1        # unsupported shape: loud, no fallback
    assert lib.nrf_mlp_small_param_count(ctypes.byref(ops.mlp_shape())) == 9344
    assert lib.nrf_composite_fwd(None, 3, None, None, None, 0.0, 0, 4, 8, None, None, None, None, None, None) == -1


def test_classic_nerf_training_sizes_and_validation_without_gpu():
    """Host-side entries of the classic-NeRF training ABI: scratch sizes (one record per 128-row tile, + one spare forward record)
    and loud refusal of shapes the kernels were not built for, null pointers and negative counts — no kernel is launched."""
    from nerfpp_b200 import cabi, ops
    lib = cabi.lib()
    shape, bad = ops.mlp_nerf_shape(), ops.mlp_nerf_shape(width=128)
    save_tile, grad_tile = 684032, 626688                                   # mlp_nerf_layout.cuh: kSaveTile, kGradTile
    for n, tiles in ((0, 0), (1, 1), (128, 1), (129, 2), (196608, 1536)):
        assert lib.nrf_mlp_nerf_saved_bytes(ctypes.byref(shape), n) == (tiles + 1) * save_tile
        assert lib.nrf_mlp_nerf_bwd_workspace_bytes(ctypes.byref(shape), n) == tiles * grad_tile
    assert lib.nrf_mlp_nerf_saved_bytes(ctypes.byref(bad), 128) == -1 and lib.nrf_mlp_nerf_bwd_workspace_bytes(ctypes.byref(bad), 128) == -1
    assert lib.nrf_mlp_nerf_saved_bytes(ctypes.byref(shape), -1) == -1
    assert lib.nrf_mlp_nerf_packed_bytes(ctypes.byref(shape)) > 1_200_000
    assert lib.nrf_mlp_nerf_fwd_train(ctypes.byref(bad), None, None, 4, None, None, None) == -3          # NRF_ERR_UNSUPPORTED
    assert lib.nrf_mlp_nerf_fwd_train(ctypes.byref(shape), None, None, 0, None, None, None) == 0         # empty batch: no-op
    assert lib.nrf_mlp_nerf_fwd_train(ctypes.byref(shape), None, None, 4, None, None, None) == -1        # null pointers
    assert lib.nrf_mlp_nerf_bwd(ctypes.byref(shape), None, None, None, 4, None, None, None) == -1
    assert lib.nrf_mlp_nerf_bwd(ctypes.byref(shape), None, None, None, -2, None, None, None) == -1
    assert b"negative" in lib.nrf_last_error()


def test_lerf_head_sizes_and_validation_without_gpu():
    """Host-side entries of the LeRF head ABI: blob / scratch sizes and loud refusal of shapes that are not built
    (the reference's default 768-d language dimension, src/NeRFExecutor.h:63, included) — no kernel is launched."""
    from nerfpp_b200 import cabi, ops
    lib = cabi.lib()
    shape = ops.lerf_shape()
    operand = 2 * (256 * 128 + 48 * 256 + 256 * 160 + 3 * 256 * 256)          # S0, S1 (33 -> 48), E0, G, E1 lower / upper halves, fp16
    train = 2 * (256 * 128 + 48 * 256 + 256 * 160 + 256 * 256)                  # S0, S1, E0, G once more in bf16 (the gradient chain's copy)
    # + W_e1^T fp32 + the scale of G + G itself in fp32
    assert lib.nrf_lerf_packed_bytes(ctypes.byref(shape)) == operand + train + 4 * 256 * 512 + 128 + 4 * 256 * 256
    for n, tiles in ((0, 0), (1, 1), (128, 1), (129, 2), (196608, 1536)):
        assert lib.nrf_lerf_hidden_bytes(ctypes.byref(shape), n) == tiles * 128 * 256 * 2
        # training records per 128-row tile: bf16 [geo | x] (160), h1, h2 (256 each), fp16 h2 (256), 32 mask bytes per row
        assert lib.nrf_lerf_train_saved_bytes(ctypes.byref(shape), n) == tiles * (128 * 2 * (160 + 3 * 256) + 128 * 32)
        assert lib.nrf_lerf_bwd_workspace_bytes(ctypes.byref(shape), n, max(n // 192, 0)) >= tiles * 128 * 2 * (3 * 256 + 48)
    for bad in (ops.lerf_shape(lang_embed_dim=768), ops.lerf_shape(num_layers=3, hidden_dim=64), ops.lerf_shape(input_ch=32)):
        assert lib.nrf_lerf_packed_bytes(ctypes.byref(bad)) == -1 and lib.nrf_lerf_hidden_bytes(ctypes.byref(bad), 128) == -1
        assert lib.nrf_lerf_fwd(ctypes.byref(bad), None, None, None, 4, None, None) == -3                 # NRF_ERR_UNSUPPORTED
    assert lib.nrf_lerf_hidden_bytes(ctypes.byref(shape), -1) == -1
    assert lib.nrf_lerf_sigma_fwd(ctypes.byref(shape), None, None, None, 0, None, None) == 0             # empty batch: no-op
    assert lib.nrf_lerf_fwd(ctypes.byref(shape), None, None, None, 4, None, None) == -1                  # null pointers
    assert lib.nrf_lerf_hidden_fwd(ctypes.byref(shape), None, None, None, 4, None, None, None, None) == -1
    assert lib.nrf_lerf_render_embedding(ctypes.byref(shape), None, None, None, None, 4, 0, None, None, None) == -1
    assert lib.nrf_lerf_render_embedding(ctypes.byref(shape), None, None, None, None, 0, 8, None, None, None) == 0
    with pytest.raises(cabi.NrfError):
        import torch
        ops.lerf_fwd(torch.zeros(8, dtype=torch.uint8), torch.zeros(4, 128, dtype=torch.float16))        # CPU tensors: no CPU path


def test_product_path_refuses_cpu_tensors():
    import torch
    from nerfpp_b200 import cabi, ops
    with pytest.raises(cabi.NrfError):
        ops.sh_encode(torch.zeros(4, 3), 4)
    with pytest.raises(cabi.NrfError):
        ops.composite_fwd(torch.zeros(2, 4, 4), torch.zeros(2, 4), torch.zeros(2, 3))


def test_product_never_imports_the_oracle():
    for p in (ROOT / "nerfpp_b200").rglob("*"):
        if p.suffix in (".py", ".cu", ".cuh", ".cpp", ".h") and p.is_file():
            text = p.read_text()
            assert not re.search(r"(import|from)\s+(oracle|restate)\b|oracle/_ref|nerfpp_ref|#include\s+\"[^\"]*oracle", text), p


def test_reference_call_sites_compile_against_the_compat_headers():
    """The statements of the reference's NeRFExecutor / main.cpp that touch the hot-path classes (tests/compat_call_sites.cpp, each citing its
    line), written against the reference's own header names, compile with nerfpp_b200/host/compat first on the include path — HashNeRF + LeRF
    and classic NeRF instantiations (syntax + template instantiation only: nothing is linked or run)."""
    import subprocess
    import sysconfig
    import torch
    tdir = Path(torch.__file__).resolve().parent
    inc = []
    for d in (ROOT / "nerfpp_b200/host/compat", ROOT / "nerfpp_b200/host", ROOT / "include", tdir / "include", tdir / "include/torch/csrc/api/include",
              Path("/usr/local/cuda/include"), Path(sysconfig.get_paths()["include"])):
        inc += ["-I", str(d)]
    r = subprocess.run(["g++", "-std=c++17", "-fsyntax-only", "-w", "-D_GLIBCXX_USE_CXX11_ABI=1", *inc, str(ROOT / "tests/compat_call_sites.cpp")],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-4000:]


def test_cpp_host_layer_loads_and_refuses_cpu_tensors():
    """nerfpp_b200_torch.so (the torch::Tensor drop-in layer) builds, imports and exposes the reference's class surface; like
    the C ABI it has no CPU path."""
    import sys
    import torch
    from nerfpp_b200 import build
    sys.path.insert(0, str(build.build_host().parent))
    import nerfpp_b200_torch as H
    for name in ("make_cuhash", "make_classic", "sample_pdf", "get_rays", "intersect_aabb", "trunc_exp", "cu_sh_encoder", "embedder"):
        assert hasattr(H, name)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        H.embedder(torch.rand(4, 3), 10)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        H.sample_pdf(torch.rand(2, 5), torch.rand(2, 4), 8, True)
    x = torch.randn(5, requires_grad=True)                       # TruncExp is plain ATen (src/CustomOps.cpp:5-16)
    y = H.trunc_exp(x * 4)
    y.sum().backward()
    assert torch.allclose(y, torch.exp(x.detach() * 4))
    assert torch.allclose(x.grad, 4 * torch.exp(torch.clamp(x.detach() * 4, -100, 5)))
