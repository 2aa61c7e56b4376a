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


# ---------------------------------------------------------------------------------------------------------------- training
# nrf_mlp_nerf_fwd_train / nrf_mlp_nerf_bwd against fp64 autograd of NeRFImpl::forward (the same graph as the pinned restatement
# oracle/restate.py:nerf_forward, with the intermediate tensors kept).  bf16 operands, fp32 accumulation.  Two references:
#   * emulate=True  — weights, inputs and every stored activation rounded to bf16 (straight-through), i.e. the kernels' rounding points,
#     so the ReLU masks are the kernels' masks: tolerance 1e-2 of sum|terms| of each gradient (north star: "1e-2 bf16");
#   * emulate=False — exact fp64.  A bf16 forward flips the ReLU mask of the ~0.3 % of units whose pre-activation is within bf16
#     rounding of zero; with a random upstream gradient those flips are incoherent noise of order sqrt(flipped / n) in every weight
#     gradient (5-15 % at n = 1000), so this comparison is made on direction (cosine), not element-wise.


def _train_reference(x, p, gout, emulate):
    return O.nerf_train_reference(x, p, gout, emulate)


def _xavier_params(seed=0, gain=0.1, w=256):
    """Trainable::Initialize (src/LibTorchTraining/Trainable.h:32-57): xavier_normal_(gain 0.1) weights, zero biases."""
    g = torch.Generator().manual_seed(seed)
    p = _params(seed, w)
    for k, v in p.items():
        if k.endswith(".weight"):
            o, i = v.shape
            p[k] = (torch.randn(o, i, generator=g) * gain * math.sqrt(2.0 / (i + o))).cuda()
        else:
            p[k] = torch.zeros_like(v)
    return p


def _check_forward(out, ref):
    """bf16 forward (weights, inputs and 10 stored activations rounded to 8 mantissa bits) against exact fp64: rel 1e-2-class in the rms
    sense (the north star's bf16 figure; measured 0.6-1.3e-2, asserted at 1.5e-2); the worst single output may sit a few sigma out.  A bf16 emulation is no tighter a
    reference here: one rounding tie broken differently (fp32 vs fp64 accumulation) re-draws the rounding noise of every later layer."""
    err = out.double() - ref
    assert float(err.pow(2).mean().sqrt()) <= 1.5e-2 * float(ref.pow(2).mean().sqrt())
    assert float(err.abs().max()) <= 4e-2 * float(ref.abs().max())


def _check_grads(grads, ref, times=1.0, tol=1e-2):
    for k, g in grads.items():
        err = float((g.double() - times * ref["grads"][k]).abs().max())
        assert err <= tol * times * ref["bound"][k], (k, err, ref["bound"][k])


def _cosines(grads, ref):
    return {k: float((g.double() * ref["grads"][k]).sum() / (g.double().norm() * ref["grads"][k].norm()).clamp_min(1e-300)) for k, g in grads.items()}


@pytest.mark.parametrize("n", [1, 127, 129, 1000, 20011])
def test_train_forward_and_backward_match_fp64_autograd(n):
    from nerfpp_b200 import ops
    p = _params(seed=n + 7)
    x = _inputs(n, seed=n + 8)
    gout = (torch.randn(n, 4, generator=torch.Generator().manual_seed(n)) * 1e-3).cuda()
    emu, exact = _train_reference(x, p, gout, True), _train_reference(x, p, gout, False)
    packed = ops.mlp_nerf_pack(p, train=True)
    out, saved = ops.mlp_nerf_fwd_train(packed, x)
    _check_forward(out, exact["out"])
    grads = ops.mlp_nerf_bwd(packed, saved, gout, {k: torch.zeros_like(v) for k, v in p.items()})
    _check_grads(grads, emu)
    if n >= 1000:
        cos = _cosines(grads, exact)
        assert min(cos.values()) >= 0.97, cos
    # += semantics: a second call doubles the result (fp32 atomics: order-dependent rounding only)
    ops.mlp_nerf_bwd(packed, saved, gout, grads)
    _check_grads(grads, emu, times=2.0)


def test_backward_survives_reference_initialisation():
    """Xavier(0.1): activations shrink ~14x per layer and the gradients of the first layers are ~1e-8 of the last layer's — below
    fp16's range, inside bf16's.  Every gradient must match relative to ITS OWN scale."""
    from nerfpp_b200 import ops
    n = 4096
    p = _xavier_params(seed=2)
    x = _inputs(n, seed=3)
    gout = torch.randn(n, 4, generator=torch.Generator().manual_seed(4)).cuda() / n
    emu, exact = _train_reference(x, p, gout, True), _train_reference(x, p, gout, False)
    packed = ops.mlp_nerf_pack(p, train=True)
    out, saved = ops.mlp_nerf_fwd_train(packed, x)
    _check_forward(out, exact["out"])
    grads = ops.mlp_nerf_bwd(packed, saved, gout, {k: torch.zeros_like(v) for k, v in p.items()})
    assert all(float(g.abs().max()) > 0 for g in grads.values())
    _check_grads(grads, emu)
    cos = _cosines(grads, exact)
    assert min(cos.values()) >= 0.97, cos


def test_training_empty_and_masked_rows():
    from nerfpp_b200 import ops
    p = _params(seed=1)
    packed = ops.mlp_nerf_pack(p, train=True)
    out, saved = ops.mlp_nerf_fwd_train(packed, torch.empty(0, 90, device="cuda"))
    assert out.shape == (0, 4)
    g0 = {k: torch.zeros_like(v) for k, v in p.items()}
    ops.mlp_nerf_bwd(packed, saved, torch.empty(0, 4, device="cuda"), g0)
    assert all(float(v.abs().max()) == 0 for v in g0.values())
    # rows with zero upstream gradient contribute nothing: gradient of the full batch == gradient of the rows that have one
    n = 700
    x = _inputs(n, seed=9)
    gout = torch.randn(n, 4, generator=torch.Generator().manual_seed(5)).cuda() * 1e-3
    gout[300:] = 0
    out, saved = ops.mlp_nerf_fwd_train(packed, x)
    full = ops.mlp_nerf_bwd(packed, saved, gout, {k: torch.zeros_like(v) for k, v in p.items()})
    out2, saved2 = ops.mlp_nerf_fwd_train(packed, x[:300].contiguous())
    part = ops.mlp_nerf_bwd(packed, saved2, gout[:300].contiguous(), {k: torch.zeros_like(v) for k, v in p.items()})
    for k in full:
        s = float(part[k].abs().max())
        assert float((full[k] - part[k]).abs().max()) <= 1e-5 * s + 1e-12, k


def test_full_size_training_batch_backward_properties():
    """BASELINE C1 training batch (1024 rays x 192 samples = 196 608 rows), size-independent properties of the backward:
    (a) linear in the upstream gradient for fixed saved activations: grads(a*g1 + g2) = a*grads(g1) + grads(g2);
    (b) a sum over rows: the gradient of the full batch equals the sum of the gradients of its two halves (each half run as its own
        forward + backward: different tiles, different work split over the SMs, different accumulation order)."""
    from nerfpp_b200 import ops
    n = 1024 * 192
    p = _params(seed=11)
    x = _inputs(n, seed=12)
    gen = torch.Generator().manual_seed(13)
    g1, g2 = (torch.randn(n, 4, generator=gen) * 1e-3).cuda(), (torch.randn(n, 4, generator=gen) * 1e-3).cuda()
    packed = ops.mlp_nerf_pack(p, train=True)
    out, saved = ops.mlp_nerf_fwd_train(packed, x)
    assert torch.isfinite(out).all()
    zeros = lambda: {k: torch.zeros_like(v) for k, v in p.items()}  # noqa: E731
    ga, gb = ops.mlp_nerf_bwd(packed, saved, g1, zeros()), ops.mlp_nerf_bwd(packed, saved, g2, zeros())
    gc = ops.mlp_nerf_bwd(packed, saved, (2.0 * g1 + g2).contiguous(), zeros())
    for k in p:
        ref = 2.0 * ga[k].double() + gb[k].double()
        assert float((gc[k].double() - ref).abs().max()) <= 1e-2 * float(ref.abs().max()) + 1e-12, k    # bf16 rounding of the gradient rows
    h = n // 2 + 64          # not a multiple of the 128-row tile on purpose
    parts = zeros()
    for lo, hi in ((0, h), (h, n)):
        o2, s2 = ops.mlp_nerf_fwd_train(packed, x[lo:hi].contiguous())
        assert torch.equal(o2, out[lo:hi])                                 # rows are independent, bit for bit
        ops.mlp_nerf_bwd(packed, s2, g1[lo:hi].contiguous(), parts)      # += into the same tensors
    for k in p:
        assert float((parts[k] - ga[k]).abs().max()) <= 1e-4 * float(ga[k].abs().max()) + 1e-12, k   # fp32 summation order only


# ---------------------------------------------------------------------------------------------------------------- fused embedding
@pytest.mark.parametrize("n_rays,s", [(1, 1), (3, 43), (100, 192), (257, 64)])
def test_points_entry_equals_embed_concat_forward_bit_for_bit(n_rays, s):
    """nrf_mlp_nerf_fwd_points / _fwd_train_points == nrf_posenc_fwd (points) + nrf_posenc_fwd (directions expanded per sample,
    src/NeRFRenderer.h:179-181) + cat (:182) + nrf_mlp_nerf_fwd / _fwd_train: the same operands, so the same bits; and the training
    records they leave give the same gradients."""
    from nerfpp_b200 import ops
    g = torch.Generator().manual_seed(n_rays * 1000 + s)
    n = n_rays * s
    pts = (torch.rand(n, 3, generator=g) * 3 - 1.5).cuda()
    dirs = torch.nn.functional.normalize(torch.randn(n_rays, 3, generator=g), dim=-1).cuda()
    fp, fv = O.posenc_freqs(10), O.posenc_freqs(4)
    x = torch.cat([ops.posenc(pts, fp), ops.posenc(dirs.repeat_interleave(s, 0).contiguous(), fv)], -1).contiguous()
    p = _params(seed=s)
    packed, packed_t = ops.mlp_nerf_pack(p), ops.mlp_nerf_pack(p, train=True)
    assert torch.equal(ops.mlp_nerf_fwd_points(packed, pts, dirs, s, fp, fv), ops.mlp_nerf_fwd(packed, x))
    out_a, saved_a = ops.mlp_nerf_fwd_train(packed_t, x)
    out_b, saved_b = ops.mlp_nerf_fwd_points(packed_t, pts, dirs, s, fp, fv, train=True)
    assert torch.equal(out_a, out_b)
    tiles = (n + 127) // 128
    act = 647168                     # activation regions of a record (kSaveBits); the mask slots behind them are only partly written (hv: 16 of 32 B)
    ra, rb = saved_a[:tiles * 684032].view(tiles, 684032), saved_b[:tiles * 684032].view(tiles, 684032)
    assert torch.equal(ra[:, :act], rb[:, :act])                                       # every stored layer input, bit for bit
    gout = (torch.randn(n, 4, generator=g) * 1e-3).cuda()
    ga = ops.mlp_nerf_bwd(packed_t, saved_a, gout, {k: torch.zeros_like(v) for k, v in p.items()})
    gb = ops.mlp_nerf_bwd(packed_t, saved_b, gout, {k: torch.zeros_like(v) for k, v in p.items()})
    for k in p:
        assert torch.allclose(ga[k], gb[k], rtol=1e-4, atol=1e-6 * float(ga[k].abs().max()) + 1e-12), k      # fp32 atomic order only
    # fused embedding vs the oracle's embedding (restatement of EmbedderImpl::forward): same values up to sinf/cosf rounding
    ref = O.nerf_forward(torch.cat([O.posenc(pts.cpu(), 10), O.posenc(dirs.cpu().repeat_interleave(s, 0), 4)], -1).double().cuda(),
                         {k: v.double() for k, v in p.items()})
    assert float((ops.mlp_nerf_fwd_points(packed, pts, dirs, s, fp, fv).double() - ref).abs().max()) <= 1e-2 * float(ref.abs().max())
