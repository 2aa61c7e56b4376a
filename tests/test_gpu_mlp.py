"""GPU parity of the fused NeRFSmall kernels (nrf_mlp_small_fwd / _bwd) against the reference's NeRFSmallImpl
(golden fixture generated from src/NeRF.cpp:363-412 through autograd) and oracle/restate.py.

Tolerance class: the reference MLP is true fp32 SGEMM; the fused kernels run bf16 x bf16 (layer 0: fp16 x fp16) with fp32
accumulation, so outputs / gradients are held to rel 1e-2 of the tensor's scale (BASELINE north star: "1e-2 bf16")."""
import numpy as np
import pytest
import torch

import restate as O
import mlp_emul as E

pytestmark = pytest.mark.gpu


def rel_err(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def flat(ws):
    return torch.cat([w.reshape(-1) for w in ws]).contiguous()


def run_fwd_bwd_f32cat(ws, x, g):
    from nerfpp_b200 import ops
    params = flat(ws).cuda()
    packed = ops.mlp_small_pack(params)
    xc = x.cuda().contiguous()
    out = ops.mlp_small_fwd(packed, xc, None, 1, None)
    gp = torch.zeros_like(params)
    gx = ops.mlp_small_bwd(packed, xc, None, 1, None, g.cuda().contiguous(), gp)
    return out, gx, gp


def check_against_emulation_and_fp32(ws, x, gout, out, gx, gp, keep=None):
    """The four-step criterion of tests/mlp_emul.py."""
    e_out, e_gx, e_dw, e_masks = E.forward_backward(ws, x, gout, keep)
    nx = gx.shape[1]
    # 1. kernel == quantisation-point-exact emulation
    assert rel_err(out, e_out) < 2e-3
    assert rel_err(gx, e_gx[:, :nx]) < 4e-3
    for i, (a, b) in enumerate(zip(split(gp, ws), e_dw)):
        assert rel_err(a, b) < 4e-3, f"dW{i} vs emulation"
    # 2. forward vs the fp32 reference (src/NeRF.cpp:363-412; sigma masked as NeRFRenderer.h:188)
    wr = [w.clone().requires_grad_(True) for w in ws]
    xr = x.clone().requires_grad_(True)
    ref = O.nerf_small_forward(xr, (wr[:2], wr[2:]))
    if keep is not None:
        ref = torch.cat([ref[:, :3], ref[:, 3:] * keep.float()[:, None]], -1)
    ref.backward(gout)
    assert rel_err(out, ref.detach()) < 1e-2
    # 3. gradients vs fp32 arithmetic on the same ReLU active sets
    r_masks = E.fp32_masks(ws, x)
    chk_gx, chk_dw = E.fp32_with_masks(ws, x, gout, r_masks, keep)     # with its own masks this IS the autograd reference
    assert rel_err(chk_gx, xr.grad.detach()) < 1e-5
    a_gx, a_dw = E.fp32_with_masks(ws, x, gout, e_masks, keep)
    assert rel_err(gx, a_gx[:, :nx]) < 1.5e-2
    for i, (a, b) in enumerate(zip(split(gp, ws), a_dw)):
        assert rel_err(a, b) < 1.5e-2, f"dW{i} vs fp32 on the same active sets"
    # 4. the active sets themselves, and the typical row against the true fp32 gradient
    flips = sum(int((a != b).sum()) for a, b in zip(e_masks, r_masks)) / sum(m.numel() for m in r_masks)
    row = (gx.double() - xr.grad[:, :nx].double()).norm(dim=1) / xr.grad[:, :nx].double().norm(dim=1).clamp_min(1e-30)
    print(f"ReLU units on the other branch than fp32: {flips:.3%}; per-row dX error median {float(row.median()):.2e} max {float(row.max()):.2e}")
    assert flips < 2e-2
    assert float(row.median()) < 1e-2
    return ref.detach()


def split(flat_grad, ws):
    out, off = [], 0
    for w in ws:
        out.append(flat_grad[off:off + w.numel()].reshape(w.shape))
        off += w.numel()
    return out


@pytest.mark.parametrize("which", ["xavier", "unit"])
def test_against_reference_fixture(golden, which):
    g = golden("nerf_small.npz")
    wk, xk, ok, gxk, gwk = ("w", "x", "out", "gx", "gw") if which == "xavier" else ("v", "x2", "out2", "gx2", "gv")
    ws = [torch.from_numpy(g[f"{wk}{i}"]) for i in range(5)]
    x = torch.from_numpy(g[xk])
    x = torch.cat([x[:, :32].half().float(), x[:, 32:]], -1)      # the kernel consumes the encodings as fp16 values
    gout = torch.from_numpy(g["g"])
    out, gx, gp = run_fwd_bwd_f32cat(ws, x, gout)
    ref = check_against_emulation_and_fp32(ws, x, gout, out.cpu(), gx.cpu(), gp.cpu())
    if which == "unit":
        assert rel_err(ref, torch.from_numpy(g[ok])) < 1e-5       # the restatement IS the reference here


def test_fused_input_path_matches_cat_path():
    """ENC16 + per-ray SH input (the pipeline path) == fp32 cat input (the drop-in NeRFSmall::forward path)."""
    from nerfpp_b200 import ops
    torch.manual_seed(3)
    rays, s = 37, 5                                            # 185 rows: ragged vs the 16-row slabs and 128-row tiles
    n = rays * s
    ws = [torch.randn(o, i) * (2.0 / i) ** 0.5 for o, i in ((64, 32), (16, 64), (64, 31), (64, 64), (3, 64))]
    params = flat(ws).cuda()
    packed = ops.mlp_small_pack(params)
    enc = torch.randn(n, 32).half().cuda()
    dirs = torch.nn.functional.normalize(torch.randn(rays, 3), dim=-1).cuda()
    ray_sh = ops.sh_encode(dirs, 4)
    keep = (torch.rand(n) > 0.2).to(torch.uint8).cuda()
    x = torch.cat([enc.float(), ray_sh.repeat_interleave(s, 0)], -1).contiguous()
    out_a = ops.mlp_small_fwd(packed, enc, ray_sh, s, keep)
    out_b = ops.mlp_small_fwd(packed, x, None, 1, keep)
    assert torch.equal(out_a, out_b)
    assert float(out_a[keep == 0, 3].abs().max()) == 0.0       # sigma := 0 outside the box (NeRFRenderer.h:188)
    g = torch.randn(n, 4).cuda()
    gp_a, gp_b = torch.zeros_like(params), torch.zeros_like(params)
    gx_a = ops.mlp_small_bwd(packed, enc, ray_sh, s, keep, g, gp_a)
    gx_b = ops.mlp_small_bwd(packed, x, None, 1, keep, g, gp_b)
    assert rel_err(gx_a.float(), gx_b[:, :32]) < 1e-2          # bf16 vs fp32 output of the same values
    assert rel_err(gp_a, gp_b) < 1e-5
    # reference semantics of the mask: gradient of sigma is dropped where keep == 0
    check_against_emulation_and_fp32(ws, x.cpu(), g.cpu(), out_b.cpu(), gx_b.cpu(), gp_b.cpu(), keep.cpu())


@pytest.mark.parametrize("degree", [4, 8, 3])
def test_per_ray_view_term_form(degree):
    """NRF_MLP_IN_ENC16_RAYBIAS: the view channels (any SH degree; 8 = the reference's shipped 64-d input, src/main.cpp:176) enter as a per-ray
    bias of the colour net's first layer.  Forward against the fp32 restatement of NeRFSmallImpl::forward at that width; gradients against
    autograd of the same; at degree 4 also against the kernels' own per-sample view input (same weights, same blob)."""
    from nerfpp_b200 import ops
    torch.manual_seed(10 + degree)
    V = degree * degree
    rays, s = 23, 32                                           # 736 rows: ragged vs the 128-row tiles; a 16-row slab lies within one ray
    n = rays * s
    shape = ops.mlp_shape(input_ch_views=V)
    ws = [torch.randn(o, i) * (2.0 / i) ** 0.5 for o, i in ((64, 32), (16, 64), (64, V + 15), (64, 64), (3, 64))]
    params = flat(ws).cuda()
    assert params.numel() == 9344 + 64 * (V - 16)
    packed = ops.mlp_small_pack(params, shape=shape)
    enc = torch.randn(n, 32).half().cuda()
    ray_sh = ops.sh_encode(torch.nn.functional.normalize(torch.randn(rays, 3), dim=-1).cuda(), degree)
    keep = (torch.rand(n) > 0.2).to(torch.uint8).cuda()
    gb = torch.full((rays, 64), 7.0, device="cuda")
    bias = ops.mlp_small_view_bias_fwd(packed, ray_sh, shape=shape, grad_bias_zero=gb)
    assert float(gb.abs().max()) == 0.0                        # the accumulator is cleared by the same launch
    assert rel_err(bias, ray_sh.cpu() @ ws[2][:, :V].t()) < 1e-5
    out = ops.mlp_small_fwd(packed, enc, bias, s, keep, shape=shape, ray_bias=True)
    # fp32 reference
    x = torch.cat([enc.float().cpu(), ray_sh.cpu().repeat_interleave(s, 0)], -1)
    wr = [w.clone().requires_grad_(True) for w in ws]
    xr = x.clone().requires_grad_(True)
    ref = O.nerf_small_forward(xr, (wr[:2], wr[2:]), input_ch_views=V)
    ref = torch.cat([ref[:, :3], ref[:, 3:] * keep.cpu().float()[:, None]], -1)
    assert rel_err(out, ref.detach()) < 1e-2
    assert float(out[keep == 0, 3].abs().max()) == 0.0
    g = torch.randn(n, 4)
    ref.backward(g)
    gp = torch.zeros_like(params)
    gx = ops.mlp_small_bwd(packed, enc, bias, s, keep, g.cuda(), gp, shape=shape, grad_bias=gb)
    ops.mlp_small_view_bias_bwd(ray_sh, gb, gp, shape=shape)

    def cos(a, b):
        a, b = a.double().flatten().cpu(), b.double().flatten().cpu()
        return float(a @ b / (a.norm() * b.norm()).clamp_min(1e-300))

    # bf16 gradient chain + a fraction of a percent of ReLU units on the other branch than fp32: a row with a flipped unit is off by that unit's
    # whole contribution, so the max-norm is loose and direction + the typical row are tight (test_against_reference_fixture separates the two
    # effects for the 16-channel form; at degree 4 the comparison below against that form is the tight one)
    row = (gx.double().cpu() - xr.grad[:, :32].double()).norm(dim=1) / xr.grad[:, :32].double().norm(dim=1).clamp_min(1e-30)
    assert cos(gx.float(), xr.grad[:, :32]) > 0.999 and float(row.median()) < 1e-2 and rel_err(gx.float(), xr.grad[:, :32]) < 0.25
    for i, (a, b) in enumerate(zip(split(gp.cpu(), ws), [w.grad for w in wr])):
        assert cos(a, b) > 0.999 and rel_err(a, b) < 0.1, f"dW{i}"
    # the view columns (per-ray route) on their own
    assert cos(split(gp.cpu(), ws)[2][:, :V], wr[2].grad[:, :V]) > 0.999
    if degree == 4:
        out_v = ops.mlp_small_fwd(packed, enc, ray_sh, s, keep)
        assert rel_err(out, out_v) < 3e-3                      # fp32 view term vs fp16 view operands in the MMA
        gp_v = torch.zeros_like(params)
        gx_v = ops.mlp_small_bwd(packed, enc, ray_sh, s, keep, g.cuda(), gp_v)
        # same weights, same blob; the two forms round the view term differently (fp32 bias vs fp16 operands inside the MMA), which moves a few
        # pre-activations across zero: direction and the typical row are tight, single rows are not
        row_v = (gx.double() - gx_v.double()).norm(dim=1) / gx_v.double().norm(dim=1).clamp_min(1e-30)
        assert cos(gx.float(), gx_v.float()) > 0.9999 and float(row_v.median()) < 5e-3 and rel_err(gx.float(), gx_v.float()) < 0.1
        assert cos(gp, gp_v) > 0.9999 and rel_err(gp, gp_v) < 5e-2
    else:
        with pytest.raises(Exception):                         # the per-sample view input is built for 16 channels only
            ops.mlp_small_fwd(packed, enc, ray_sh, s, keep, shape=shape)
    # importance-only forward on the same form: rows at perm positions
    N, S = 16, 16
    perm = torch.stack([torch.randperm(s)[:s] for _ in range(rays)]).to(torch.int16).cuda()      # [R, T = 32]: first N entries = new rows
    raw_m = torch.zeros(n, 4, device="cuda")
    ops.mlp_small_fwd_importance(packed, enc, bias, keep, perm, N, raw_m, shape=shape, ray_bias=True)
    rows = (torch.arange(rays, device="cuda")[:, None] * s + perm[:, :N].long()).flatten()
    assert torch.equal(raw_m[rows], out[rows])
    untouched = torch.ones(n, dtype=torch.bool, device="cuda")
    untouched[rows] = False
    assert float(raw_m[untouched].abs().max()) == 0.0


def test_sizes_and_accumulation():
    from nerfpp_b200 import ops
    torch.manual_seed(0)
    ws = [torch.randn(o, i) * (2.0 / i) ** 0.5 for o, i in ((64, 32), (16, 64), (64, 31), (64, 64), (3, 64))]
    params = flat(ws).cuda()
    packed = ops.mlp_small_pack(params)
    for n in (0, 1, 15, 16, 17, 127, 128, 129, 1000):
        x = torch.cat([torch.randn(n, 32).half().float(), torch.randn(n, 16)], -1).cuda()
        out = ops.mlp_small_fwd(packed, x, None, 1, None)
        assert out.shape == (n, 4)
        if n:
            ref = O.nerf_small_forward(x.cpu(), (ws[:2], ws[2:]))
            assert rel_err(out, ref) < 1e-2, n
    # the parameter gradient ACCUMULATES across calls (the coarse/fine passes of one step share it)
    x = torch.cat([torch.randn(300, 32).half().float(), torch.randn(300, 16)], -1).cuda()
    g = torch.randn(300, 4).cuda()
    gp1 = torch.zeros_like(params)
    ops.mlp_small_bwd(packed, x, None, 1, None, g, gp1)
    gp2 = gp1.clone()
    ops.mlp_small_bwd(packed, x, None, 1, None, g, gp2)
    assert rel_err(gp2, 2 * gp1) < 1e-5


def test_unsupported_shape_fails_loudly():
    from nerfpp_b200 import cabi, ops
    with pytest.raises(cabi.NrfError):
        ops.mlp_small_pack(torch.zeros(9344, device="cuda"), shape=ops.mlp_shape(hidden=128))


def test_full_size_linearity():
    """BASELINE size (786 432 rows): the backward is linear in grad_raw — bwd(a*g1 + g2) == a*bwd(g1) + bwd(g2)."""
    from nerfpp_b200 import ops
    torch.manual_seed(1)
    n = 4096 * 192
    ws = [torch.randn(o, i) * (2.0 / i) ** 0.5 for o, i in ((64, 32), (16, 64), (64, 31), (64, 64), (3, 64))]
    params = flat(ws).cuda()
    packed = ops.mlp_small_pack(params)
    enc = torch.randn(n, 32, device="cuda").half()
    ray_sh = ops.sh_encode(torch.nn.functional.normalize(torch.randn(4096, 3, device="cuda"), dim=-1), 4)
    g1, g2 = torch.randn(n, 4, device="cuda"), torch.randn(n, 4, device="cuda")
    outs = []
    for g in (g1, g2, 0.5 * g1 + g2):
        gp = torch.zeros_like(params)
        ops.mlp_small_bwd(packed, enc, ray_sh, 192, None, g, gp, want_grad_in=False)
        outs.append(gp)
    assert rel_err(outs[2], 0.5 * outs[0] + outs[1]) < 2e-2
