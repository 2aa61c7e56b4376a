"""GPU parity of the fused LeRF training path (nrf_lerf_fwd_train, nrf_lerf_render_embedding_train, nrf_lerf_bwd_rays, nrf_lerf_bwd_rows) through the
C ABI against the oracle: oracle/restate.py:lerf_backward_fused_form — pinned in fp64 against autograd through the reference's own LeRF +
LeRFRenderer::RawToLEOutputs + the language loss (tests/test_oracle_pin.py, tests/golden/lerf_grads.npz).  bf16 tensor-core arithmetic: the
tolerance is the north star's bf16 figure (rel 1e-2) on the forward and on every gradient at the kernels' rounding points."""
import math

import numpy as np
import pytest
import torch

import restate as O

pytestmark = pytest.mark.gpu
NAMES = ("lang_model_sigma_le_net_0.weight", "lang_model_sigma_le_net_1.weight", "lang_model_le_net_0.weight", "lang_model_le_net_1.weight")
KEYS = ("sigma_w0", "sigma_w1", "le_w0", "le_w1")


def _run(x16, w, z, d, target, keep=None):
    """x16 [R,S,128] fp16, w: 4 fp32 weights -> (loss, out, grads (dict by KEYS), d_x [R,S,128])."""
    from nerfpp_b200 import ops
    from nerfpp_b200.lerf import head_forward_backward
    r, s, _ = x16.shape
    weights = {n: t.float().cuda().contiguous() for n, t in zip(NAMES, w)}
    grads = {n: torch.zeros_like(t) for n, t in weights.items()}
    packed = ops.lerf_pack(weights)
    loss = torch.zeros(1, device="cuda")
    out, d_enc = head_forward_backward(packed, weights, x16.reshape(-1, 128).cuda().contiguous(), keep, z.float().cuda().contiguous(),
                                       d.float().cuda().contiguous(), target.float().cuda().contiguous(), grads, loss_out=loss)
    torch.cuda.synchronize()
    return float(loss), out, {k: grads[n].double().cpu() for k, n in zip(KEYS, NAMES)}, d_enc.float().double().cpu().reshape(r, s, 128)


def _rel(a, b):
    return float((a - b).norm() / b.norm().clamp_min(1e-300))


def _case(r, s, seed, sigma_gain=4.0):
    g = torch.Generator().manual_seed(seed)
    w = [torch.randn(o, i, generator=g) * math.sqrt(2.0 / i) for o, i in ((256, 128), (33, 256), (256, 160), (512, 256))]
    w[1][0] *= sigma_gain                                        # densities of O(1): rays terminate inside the interval
    x = (torch.randn(r, s, 128, generator=g) * 0.5).half()
    z = 2 + torch.sort(torch.rand(r, s, generator=g) * 4, -1).values
    d = torch.randn(r, 3, generator=g)
    target = torch.nn.functional.normalize(torch.randn(r, 512, generator=g), dim=-1)
    return w, x, z, d, target


@pytest.mark.parametrize("r,s,seed", [(64, 48, 0), (37, 192, 1), (5, 24, 2)])
def test_fused_training_forward_and_backward_match_the_oracle(r, s, seed):
    """Two references, both the fp64 oracle (oracle/restate.py:lerf_backward_fused_form) on the same fp16 encodings:
      * emulate=True  — with the kernels' rounding points (fp16 operands and stored activations in the forward, bf16 operands and gradient rows in
        the chain), so the ReLU active sets are the kernels' own: rel 1e-2 norm-wise on every gradient (the north star's bf16 figure);
      * emulate=False — exact.  Rounding h1 / geo flips the h2 units whose pre-activation is within rounding of zero; each flipped unit is an O(1)
        error of its own gradient, i.e. noise of order sqrt(flipped share), and the density gradient is a cancellation-prone residual of the
        compositing backward — a direction + scale check (cosine / norm-wise 1e-1), as for the classic MLP (tests/test_gpu_mlp_nerf.py).
    Ragged tile counts (rows % 128 != 0) included."""
    w, x, z, d, target = _case(r, s, seed)
    loss, out, grads, d_x = _run(x, w, z, d, target)
    wd = [t.double() for t in w]
    ref = O.lerf_backward_fused_form(x.double(), wd[:2], wd[2:], z.double(), d.double(), target.double())
    emu = O.lerf_backward_fused_form(x.double(), wd[:2], wd[2:], z.double(), d.double(), target.double(), emulate=True)
    ref_loss = float(O.lerf_language_loss(ref["rendered"], target.double()))
    cos = torch.nn.functional.cosine_similarity(out["rendered"].double().cpu(), ref["rendered"], dim=-1)
    print(f"R {r} S {s}: loss {loss:.6f} vs {ref_loss:.6f}; rendered cosine min {float(cos.min()):.6f}\n   vs exact fp64:      "
          + ", ".join(f"{k} {_rel(grads[k], ref[k]):.2e}" for k in KEYS) + f", x {_rel(d_x, ref['x']):.2e}\n   vs rounding points: "
          + ", ".join(f"{k} {_rel(grads[k], emu[k]):.2e}" for k in KEYS) + f", x {_rel(d_x, emu['x']):.2e}")
    assert abs(loss - ref_loss) <= 1e-2 * abs(ref_loss)
    assert float(cos.min()) > 1 - 1e-3
    assert float((out["rendered"].double().cpu() - ref["rendered"]).abs().max()) <= 1e-2 * float(ref["rendered"].abs().max())
    for k in KEYS:
        assert _rel(grads[k], emu[k]) <= 1e-2, (k, _rel(grads[k], emu[k]))
        assert float((grads[k] * ref[k]).sum() / (grads[k].norm() * ref[k].norm())) >= 0.995 and _rel(grads[k], ref[k]) <= 1e-1, k
    assert _rel(d_x, emu["x"]) <= 1e-2
    assert float((d_x * ref["x"]).sum() / (d_x.norm() * ref["x"].norm())) >= 0.995 and _rel(d_x, ref["x"]) <= 1e-1


def test_against_the_reference_autograd_fixture(golden):
    """tests/golden/lerf_grads.npz: gradients of the REFERENCE's LeRF + RawToLEOutputs + language loss under LibTorch autograd (fp64 run of the
    compiled reference, tests/golden/make_golden.py) on the lerf.npz inputs; 96 rows only, so a flipped ReLU unit is visible: 5e-2 norm-wise."""
    g, f = golden("lerf.npz"), golden("lerf_grads.npz")
    T = lambda a: torch.from_numpy(np.asarray(a))  # noqa: E731
    w = [T(g["sw0"]).float(), T(g["sw1"]).float().clone(), T(g["lw0"]).float(), T(g["lw1"]).float()]
    w[1][0] *= 4.0
    r, s = 4, 24
    x = T(g["x"]).float().reshape(r, s, 128)
    loss, out, grads, d_x = _run(x.half(), w, T(g["z"]), T(g["rays_d"]), T(f["target"]))
    assert abs(loss - float(f["loss"])) <= 1e-2 * float(f["loss"])
    assert float((out["rendered"].double().cpu() - T(f["rendered"])).abs().max()) <= 1e-2
    assert _rel(d_x, T(f["g_x"])) <= 5e-2
    for k in KEYS:
        assert abs(float(grads[k].norm()) - float(f[f"g_{k}_norm"])) <= 5e-2 * float(f[f"g_{k}_norm"]), k
        assert _rel(grads[k][:8], T(f[f"g_{k}_rows"])) <= 8e-2, (k, _rel(grads[k][:8], T(f[f"g_{k}_rows"])))


def test_keep_mask_grad_rendered_entry_and_accumulation():
    """(a) rows outside the box (keep = 0) have sigma_le = 0 and pass no density gradient; (b) an externally supplied d loss / d rendered equals the
    built-in loss path; (c) gradients accumulate (+=) over two calls."""
    from nerfpp_b200 import ops
    w, x, z, d, target = _case(16, 40, 3)
    keep = (torch.rand(16 * 40, generator=torch.Generator().manual_seed(4)) > 0.2).to(torch.uint8).cuda()
    loss, out, grads, d_x = _run(x, w, z, d, target, keep)
    assert bool((out["raw4"][keep == 0, 3] == 0).all())
    assert bool((out["d_raw4"].view(-1, 4)[keep == 0] == 0).all())
    assert bool(torch.isfinite(d_x).all()) and float(out["weights"].view(-1)[keep == 0].abs().max()) == 0.0      # no density, no weight
    # (b) + (c) through the raw entries
    weights = {n: t.float().cuda().contiguous() for n, t in zip(NAMES, w)}
    packed = ops.lerf_pack(weights)
    enc = x.reshape(-1, 128).cuda().contiguous()
    zc, dc, tc = z.float().cuda().contiguous(), d.float().cuda().contiguous(), target.float().cuda().contiguous()
    raw4, saved, q = ops.lerf_fwd_train(packed, enc, keep)
    comp = ops.composite_fwd(raw4.view(16, 40, 4), zc, dc)
    rendered, hsum, enorm = ops.lerf_render_embedding_train(packed, comp["weights"], saved, q)
    e = rendered - tc
    g_r = torch.where(e.abs() < 1.25, e, 1.25 * torch.sign(e)) / 16
    g1 = {n: torch.zeros_like(t) for n, t in weights.items()}
    ws = ops.lerf_bwd_workspace(16 * 40, 16, "cuda")
    for _ in range(2):
        dw = ops.lerf_bwd_rays(weights, saved, q, comp["weights"], hsum, rendered, enorm, g1[NAMES[3]], ws, grad_rendered=g_r.contiguous())
        d_raw4 = ops.composite_bwd(raw4.view(16, 40, 4), zc, dc, g_weights=dw)
        d_enc = ops.lerf_bwd_rows(packed, weights, saved, keep, d_raw4, 40, ws, g1)
    for k, n in zip(KEYS, NAMES):
        assert _rel(g1[n].double().cpu(), 2 * grads[k]) <= 1e-4, k
    assert _rel(d_enc.float().double().cpu().reshape(16, 40, 128), d_x) <= 1e-6


def test_full_size_c5_batch_properties():
    """BASELINE C5 batch (1024 rays x 192 fine samples = 196 608 rows): finite everywhere, every row independent of its tile (a sub-batch of
    whole rays reproduces its slice of d_enc bit for bit), and the loss decreases along the negative gradient."""
    from nerfpp_b200 import ops
    from nerfpp_b200.lerf import head_forward_backward
    w, x, z, d, target = _case(1024, 192, 7)
    weights = {n: t.float().cuda().contiguous() for n, t in zip(NAMES, w)}
    packed = ops.lerf_pack(weights)
    enc = x.reshape(-1, 128).cuda().contiguous()
    zc, dc, tc = z.float().cuda().contiguous(), d.float().cuda().contiguous(), target.float().cuda().contiguous()
    grads = {n: torch.zeros_like(t) for n, t in weights.items()}
    loss = torch.zeros(1, device="cuda")
    out, d_enc = head_forward_backward(packed, weights, enc, None, zc, dc, tc, grads, loss_out=loss)
    assert bool(torch.isfinite(d_enc.float()).all()) and all(bool(torch.isfinite(g).all()) for g in grads.values())
    lo, hi = 256, 384                                             # 128 rays = 192 tiles, tile-aligned
    g2 = {n: torch.zeros_like(t) for n, t in weights.items()}
    l2 = torch.zeros(1, device="cuda")
    out2, d2 = head_forward_backward(packed, weights, enc[lo * 192:hi * 192].contiguous(), None, zc[lo:hi].contiguous(), dc[lo:hi].contiguous(),
                                     tc[lo:hi].contiguous(), g2, loss_out=l2, grad_scale=(hi - lo) / 1024.0)
    assert torch.equal(out2["rendered"], out["rendered"][lo:hi])
    assert torch.equal(d2, d_enc[lo * 192:hi * 192])
    # a small step against the gradient lowers the loss
    step = {n: weights[n] - 0.05 * grads[n] / grads[n].norm().clamp_min(1e-20) * weights[n].norm() for n in NAMES}
    packed2 = ops.lerf_pack(step)
    l3 = torch.zeros(1, device="cuda")
    head_forward_backward(packed2, step, enc, None, zc, dc, tc, {n: torch.zeros_like(t) for n, t in weights.items()}, loss_out=l3)
    assert float(l3) < float(loss), (float(l3), float(loss))


def _field(seed=5):
    from nerfpp_b200.lerf import LeRFField
    f = LeRFField(log2_hashmap_size=14, seed=seed, lrate_decay=1, lr=2e-3)
    g = torch.Generator().manual_seed(seed)
    for v in f.weights.values():                                  # O(1) signals: the reference's U[0,1e-4) table gives ONE density sign for all points
        v.copy_((torch.randn(v.shape, generator=g) * (2.0 / v.shape[1]) ** 0.5).cuda())
    f.params[:f.n_table].copy_((torch.rand(f.n_table, generator=g) * 2 - 1).cuda())
    f.refresh()
    return f


def test_language_field_trains_eagerly_and_as_a_graph():
    """LeRFField.train_step (render + language loss + fused backward + hash scatter at F = 8 + one Adam launch over [table | weights]): the loss goes
    down, the captured graph follows the eager trajectory, the step touches table and weights, and the fine-pass row reuse changes nothing."""
    from nerfpp_b200.pipeline import synthetic_rays
    rays = 256
    o, d, _ = synthetic_rays(rays, seed=2)
    tgt = torch.nn.functional.normalize(torch.randn(rays, 512, generator=torch.Generator().manual_seed(3)), dim=-1).cuda()
    a, b, c = _field(), _field(), _field()
    c.reuse_coarse_rows = False
    p0 = a.params.clone()
    la = [float(a.train_step(o, d, tgt)) for _ in range(12)]
    b.capture_train_step(rays)
    lb = [float(b.train_step_graph(o, d, tgt)) for _ in range(12)]
    lc = [float(c.train_step(o, d, tgt)) for _ in range(3)]
    print("language loss eager", [round(v, 5) for v in la], "graph", [round(v, 5) for v in lb])
    assert all(np.isfinite(la)) and la[-1] < 0.97 * la[0], la
    np.testing.assert_allclose(lb, la, rtol=2e-2)
    np.testing.assert_allclose(lc, la[:3], rtol=1e-3)
    moved = (a.params - p0).abs()
    assert float(moved[:a.n_table].max()) > 0 and float(moved[a.n_table:].max()) > 0
    assert float(a.grads.abs().max()) == 0.0                      # cleared by the Adam kernel
    # inference on the trained field agrees with the training forward's rendering (fp16 vs bf16 operands)
    out_t = a.forward_backward(o, d, tgt)
    a.grads.zero_()
    out_i = a.render_rays(o, d)
    cos = torch.nn.functional.cosine_similarity(out_t["rendered"], out_i["rendered"], dim=-1)
    hit = out_i["acc"] > 0.5
    assert float(cos[hit].min()) > 0.995, float(cos[hit].min())
