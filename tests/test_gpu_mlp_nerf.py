"""GPU parity of the fused classic-NeRF forward (nrf_mlp_nerf_fwd, tcgen05) against oracle/restate.py:nerf_forward — the
restatement of NeRFImpl::forward (src/NeRF.cpp:92-126) that tests/test_oracle_pin.py pins against the reference's own output —
evaluated in fp32 on the same seeded weights and inputs.  Floating-point kernel: fp16 operands, fp32 accumulation; tolerance
rel 1e-2 of the output scale (the north star's bf16-class bound), and much tighter against an fp16-operand emulation."""
import math

import numpy as np
import pytest
import torch

import restate as O

pytestmark = pytest.mark.gpu


def _params(seed=0, w=256, gain=1.0):
    g = torch.Generator().manual_seed(seed)
    shapes = {f"model_pts_linears_{i}": (w, 63 if i == 0 else (w + 63 if i == 5 else w)) for i in range(8)}
    shapes.update({"model_feature_linear": (w, w), "model_alpha_linear": (1, w), "model_views_linears_0": (w // 2, w + 27),
                   "model_rgb_linear": (3, w // 2)})
    p = {}
    for name, (o, i) in shapes.items():
        p[name + ".weight"] = (torch.randn(o, i, generator=g) * gain * math.sqrt(2.0 / i)).cuda()      # He init keeps activations O(1)
        p[name + ".bias"] = (torch.randn(o, generator=g) * 0.1).cuda()
    return p


def _inputs(n, seed=1):
    g = torch.Generator().manual_seed(seed)
    pts = torch.rand(n, 3, generator=g) * 2 - 1
    dirs = torch.nn.functional.normalize(torch.randn(n, 3, generator=g), dim=-1)
    return torch.cat([O.posenc(pts, 10), O.posenc(dirs, 4)], -1).contiguous().cuda()      # [n, 63 + 27]


def _emulate(x, p):
    """The kernel's rounding points: fp16 operands (inputs, weights, every activation), fp32 accumulation and bias add."""
    hf = lambda t: t.half().float()  # noqa: E731
    q = {k: (hf(v) if k.endswith("weight") else v) for k, v in p.items()}
    pts, views = hf(x[:, :63]), hf(x[:, 63:])
    h = pts
    for i in range(8):
        h = hf(torch.relu(h @ q[f"model_pts_linears_{i}.weight"].t() + q[f"model_pts_linears_{i}.bias"]))
        if i == 4:
            h = torch.cat([pts, h], -1)
    alpha = h @ q["model_alpha_linear.weight"].t() + q["model_alpha_linear.bias"]
    feat = hf(h @ q["model_feature_linear.weight"].t() + q["model_feature_linear.bias"])
    h = hf(torch.relu(torch.cat([feat, views], -1) @ q["model_views_linears_0.weight"].t() + q["model_views_linears_0.bias"]))
    return torch.cat([h @ q["model_rgb_linear.weight"].t() + q["model_rgb_linear.bias"], alpha], -1)


@pytest.mark.parametrize("n", [1, 127, 128, 129, 1000, 20011])
def test_forward_matches_oracle(n):
    from nerfpp_b200 import ops
    p = _params(seed=n)
    x = _inputs(n, seed=n + 1)
    packed = ops.mlp_nerf_pack(p)
    out = ops.mlp_nerf_fwd(packed, x)
    ref = O.nerf_forward(x.double(), {k: v.double() for k, v in p.items()})
    scale = float(ref.abs().max())
    assert float((out.double() - ref).abs().max()) <= 1e-2 * scale, (float((out.double() - ref).abs().max()), scale)
    emu = _emulate(x.double(), {k: v.double() for k, v in p.items()})
    assert float((out.double() - emu).abs().max()) <= 2e-3 * scale        # same rounding points: only accumulation order differs


def test_empty_unsupported_and_repeatable():
    from nerfpp_b200 import cabi, ops
    p = _params()
    packed = ops.mlp_nerf_pack(p)
    assert ops.mlp_nerf_fwd(packed, torch.empty(0, 90, device="cuda")).shape == (0, 4)
    with pytest.raises(cabi.NrfError):
        ops.mlp_nerf_pack(p, shape=ops.mlp_nerf_shape(width=128))
    x = _inputs(4096 * 8)
    a, b = ops.mlp_nerf_fwd(packed, x), ops.mlp_nerf_fwd(packed, x)
    assert torch.equal(a, b)                                              # no atomics, fixed order: bit-repeatable


def test_full_size_render_batch_is_row_independent():
    """BASELINE C1 size (100x100 rays x 192 samples = 1.92 M rows): every row depends only on its own input — the result of the
    full batch equals the result of any sub-batch, bit for bit (catches tile / pipeline cross-talk at scale)."""
    from nerfpp_b200 import ops
    p = _params(seed=5)
    packed = ops.mlp_nerf_pack(p)
    n = 100 * 100 * 192
    x = _inputs(n, seed=6)
    full = ops.mlp_nerf_fwd(packed, x)
    assert torch.isfinite(full).all()
    for lo, hi in ((0, 5000), (777 * 128 + 3, 777 * 128 + 3 + 4099), (n - 130, n)):
        assert torch.equal(full[lo:hi], ops.mlp_nerf_fwd(packed, x[lo:hi].contiguous()))
