"""GPU parity of the rendering-side kernels through the C ABI: compositing fwd/bwd, inverse-CDF sampling + merge,
ray generation / AABB / z / points, SH + positional encoders, huber + Adam.  Oracles: the golden fixtures generated
from the reference (tests/golden/make_golden.py) and oracle/restate.py on the same seeded inputs."""
import numpy as np
import pytest
import torch

import restate as O

pytestmark = pytest.mark.gpu
T = torch.from_numpy


def close(a, b, rtol, atol):
    a = a.detach().cpu().numpy() if isinstance(a, torch.Tensor) else np.asarray(a)
    b = b.detach().cpu().numpy() if isinstance(b, torch.Tensor) else np.asarray(b)
    np.testing.assert_allclose(a, b, rtol=rtol, atol=atol)


# ------------------------------------------------------------------------------------------------ compositing
@pytest.mark.parametrize("white", [False, True])
def test_composite_against_reference_fixture(golden, white):
    from nerfpp_b200 import ops
    g = golden("raw_to_outputs.npz")
    tag = "w" if white else "b"
    raw, z, d = T(g["raw"]).cuda(), T(g["z"]).cuda(), T(g["rays_d"]).cuda()
    out = ops.composite_fwd(raw, z, d, white)
    for k in ("rgb", "depth", "disp", "acc", "weights"):
        close(out[k], g[f"{tag}_{k}"], rtol=1e-3, atol=1e-6)                 # rel <= 1e-3
    ups = {k: T(g[f"{tag}_g_{k}"]).cuda().contiguous() for k in ("rgb", "depth", "disp", "acc", "weights")}
    d_raw = ops.composite_bwd(raw, z, d, white, g_rgb=ups["rgb"], g_depth=ups["depth"], g_disp=ups["disp"],
                              g_acc=ups["acc"], g_weights=ups["weights"])
    ref = g[f"{tag}_d_raw"]
    close(d_raw, ref, rtol=1e-3, atol=1e-3 * np.abs(ref).max() * 1e-2)


@pytest.mark.parametrize("S", [1, 31, 64, 192, 256, 257, 448])      # > 256: the two-pass backward (the register-resident kernels stop at 256)
def test_composite_random(S):
    from nerfpp_b200 import ops
    torch.manual_seed(S)
    R = 67
    raw = torch.randn(R, S, 4) * 2
    raw[3, :, 3] = -5.0
    z = 2 + torch.sort(torch.rand(R, S) * 4, -1).values
    d = torch.randn(R, 3)
    g_rgb = torch.randn(R, 3)
    x = raw.clone().requires_grad_(True)
    ref = O.raw_to_outputs(x, z, d)
    (ref["rgb"] * g_rgb).sum().backward()
    out = ops.composite_fwd(raw.cuda(), z.cuda(), d.cuda())
    # S == 1: the reference's dists/weights degenerate to EMPTY [R,0] tensors (NeRFRenderer.h:239-240) and every map is 0
    for k in ("rgb", "depth", "disp", "acc") + (("weights",) if S > 1 else ()):
        close(out[k], ref[k], rtol=1e-3, atol=1e-6)
    if S == 1:
        assert float(out["weights"].abs().max()) == 0.0
    d_raw = ops.composite_bwd(raw.cuda(), z.cuda(), d.cuda(), g_rgb=g_rgb.cuda())
    close(d_raw, x.grad, rtol=1e-3, atol=1e-5 * float(x.grad.abs().max()))


def test_composite_noise_and_stride():
    from nerfpp_b200 import ops
    torch.manual_seed(5)
    R, S = 9, 64
    raw7 = torch.randn(R, S, 7)                               # [rgb, sigma, normals(3)] rows (use_pred_normal layout)
    z = 2 + torch.sort(torch.rand(R, S) * 4, -1).values
    d = torch.randn(R, 3)
    noise = torch.randn(R, S)
    ref = O.raw_to_outputs(raw7[..., :4], z, d, 0.7, False, noise)
    out = ops.composite_fwd(raw7.cuda(), z.cuda(), d.cuda(), False, noise.cuda(), 0.7)
    close(out["rgb"], ref["rgb"], rtol=1e-3, atol=1e-6)
    close(out["weights"], ref["weights"], rtol=1e-3, atol=1e-6)


# ------------------------------------------------------------------------------------------------ sampler
@pytest.mark.parametrize("S,N", [(64, 128), (64, 192), (33, 50)])
def test_sampler_block_kernel_equals_warp_kernel(S, N):
    """nrf_sample_pdf_merge_rows picks a CTA-per-ray kernel for small batches (<= 16384 rays: a training step) and a warp-per-ray kernel for
    large ones (a render chunk).  Same arithmetic per element: every output must agree bit for bit — depths, samples, positions and the
    travelling raw rows — including zero-weight rays and rays whose fp32 coarse depths are not monotone (they miss the box)."""
    from nerfpp_b200 import ops
    g = torch.Generator().manual_seed(S * 1000 + N)
    R = 20000                                                                 # one call: warp kernel; two halves: CTA kernel
    near = 2 + torch.rand(R, 1, generator=g)
    far = near + 3 * torch.rand(R, 1, generator=g)
    far[::97] = near[::97] + 1e-6                                             # IntersectWithAABB's far for a ray that misses the box
    t = torch.linspace(0, 1, S)
    z = (near * (1 - t) + far * t).cuda()
    w = torch.rand(R, S, generator=g) ** 4
    w[::53] = 0.0
    w[1::211, 5:] = 0.0                                                       # all the mass in a few bins: long runs of equal samples
    w = w.cuda()
    u = torch.linspace(0, 1, N).cuda()
    rows = torch.randn(R * S, 4, generator=g).cuda()
    whole = ops.sample_pdf_merge(z, w, u, want_samples=True, want_perm=True, raw_coarse=rows)
    h = R // 2
    parts = [ops.sample_pdf_merge(z[a:b].contiguous(), w[a:b].contiguous(), u, want_samples=True, want_perm=True, raw_coarse=rows[a * S:b * S].contiguous())
             for a, b in ((0, h), (h, R))]
    pos = whole[2].long()[:, N:]                                              # merged positions of the coarse samples: the only rows written
    for i, name in enumerate(("z_merged", "z_samples", "perm")):
        assert torch.equal(whole[i], torch.cat([p[i] for p in parts], 0)), name
    got = torch.cat([p[3].view(-1, S + N, 4) for p in parts], 0)
    idx = pos[:, :, None].expand(-1, -1, 4)
    assert torch.equal(torch.gather(whole[3].view(R, S + N, 4), 1, idx), torch.gather(got, 1, idx))
    assert torch.equal(torch.gather(got, 1, idx), rows.view(R, S, 4))
    assert torch.equal(torch.sort(whole[0], -1).values, whole[0])


def assert_samples_match(got, bins, w, n, ref=None):
    """got == the restatement with correctly rounded sums (tight), and that restatement == the torch-summed reference
    except at ulp ties between u and a cdf knot (count reported and bounded; see oracle/restate.py sample_pdf)."""
    exact, inds_e = O.sample_pdf(bins, w, n, True, sums="exact")
    torch_ref, inds_t = O.sample_pdf(bins, w, n, True)
    if ref is not None:                                                      # the committed fixture IS the torch-summed result
        close(torch_ref, ref, rtol=1e-6, atol=1e-6)
    close(got, exact, rtol=2e-6, atol=2e-6)
    moved = inds_e != inds_t
    wp = w + 1e-8
    cdf = torch.cat([torch.zeros_like(wp[:, :1]), torch.cumsum(wp / wp.sum(-1, keepdim=True), -1)], -1)
    u = torch.linspace(0.0, 1.0, n).expand(bins.shape[0], n)
    lo, hi = torch.minimum(inds_e, inds_t), torch.maximum(inds_e, inds_t)
    # every knot the index moved across lies within a few ulp of u: cdf[lo] and cdf[hi-1] bracket them (cdf is sorted)
    k_lo = torch.gather(cdf, -1, lo.clamp(0, cdf.shape[-1] - 1))
    k_hi = torch.gather(cdf, -1, (hi - 1).clamp(0, cdf.shape[-1] - 1))
    assert bool((((u - k_lo).abs() <= 5e-7) & ((u - k_hi).abs() <= 5e-7))[moved].all())
    same = ~moved
    # t = (u - c0) / denom amplifies the last-ulp difference of a knot by 1/denom (down to the 1e-5 guard): up to ~1e-2 of a bin
    close(got[same], torch_ref[same], rtol=1e-4, atol=2e-4)
    assert float((got[same] - torch_ref[same]).abs().median()) <= 1e-6
    print(f"searchsorted index moved at an ulp tie for {int(moved.sum())} of {moved.numel()} samples")
    assert int(moved.sum()) <= max(bins.shape[0], int(moved.numel() * 2e-3))   # at most ~the u == 1.0 sample of each ray
    assert bool((got[:, 1:] >= got[:, :-1]).all())                           # monotone: the rank merge relies on it


def test_sample_pdf_against_reference_fixture(golden):
    from nerfpp_b200 import ops
    g = golden("sample_pdf.npz")
    u = torch.linspace(0.0, 1.0, 128).cuda()
    bins, w = T(g["bins"]), T(g["weights"])
    s = ops.sample_pdf(bins.cuda(), w.cuda(), u)
    assert_samples_match(s.cpu(), bins, w, 128, T(g["samples"]))
    # the fused call site: weights[:,1:-1] and z_mid are taken inside the kernel
    w_full = torch.cat([torch.zeros(8, 1), w, torch.zeros(8, 1)], -1).cuda()
    merged, zs = ops.sample_pdf_merge(T(g["z"]).cuda(), w_full, u, want_samples=True)
    assert merged.shape == (8, 192)                                           # sample count exact
    assert torch.equal(zs, s)
    # merge == torch::sort(cat(z, z_samples)) (NeRFRenderer.h:431), bit for bit on the same samples
    assert torch.equal(merged.cpu(), torch.sort(torch.cat([T(g["z"]), zs.cpu()], -1), -1).values)
    tie_free = (zs.cpu() - T(g["samples"])).abs().max(-1).values < 1e-4
    close(merged.cpu()[tie_free], T(g["merged"])[tie_free], rtol=1e-5, atol=1e-5)
    assert int(tie_free.sum()) >= 4


def test_sample_pdf_indices_and_ties():
    """Index exactness on spiky pdfs (most bins below the denom < 1e-5 guard, where a moved index is visible)."""
    from nerfpp_b200 import ops
    torch.manual_seed(11)
    R, B, N = 512, 63, 128
    bins = torch.sort(torch.rand(R, B) * 4 + 2, -1).values
    w = torch.rand(R, B - 1) ** 6
    got = ops.sample_pdf(bins.cuda(), w.cuda(), torch.linspace(0.0, 1.0, N).cuda()).cpu()
    assert_samples_match(got, bins, w, N)
    # per-ray random u: bitonic path
    u = torch.rand(R, N)
    ref_r, _ = O.sample_pdf(bins, w, N, False, u)
    z = torch.sort(torch.rand(R, 64) * 4 + 2, -1).values
    wfull = torch.rand(R, 64) ** 4
    merged, zs = ops.sample_pdf_merge(z.cuda(), wfull.cuda(), u.cuda(), want_samples=True)
    zmid = 0.5 * (z[:, 1:] + z[:, :-1])
    ref_s, _ = O.sample_pdf(zmid, wfull[:, 1:-1], N, False, u)
    close(zs, ref_s, rtol=1e-4, atol=1e-4)
    close(merged, torch.sort(torch.cat([z, zs.cpu()], -1), -1).values, rtol=0, atol=0)


def test_merge_equals_sort_on_degenerate_rays():
    """A ray that misses the box gets far = near + 1e-6 (RayUtils.h:123): near*(1-t)+far*t is then NOT monotone in fp32 and
    every weight is 0.  The reference sorts (NeRFRenderer.h:431); the rank merge must give the same floats in the same order."""
    from nerfpp_b200 import ops
    near = torch.tensor([3.396746873855591, 2.0, 7.123456, 0.5])
    rb = torch.zeros(4, 11)
    rb[:, 6], rb[:, 7] = near, near + 1e-6
    rb[3, 7] = 2.5                                                           # one ordinary ray alongside
    t = torch.linspace(0.0, 1.0, 64)
    z = ops.z_sample(rb.cuda(), t.cuda())
    assert not bool((z[0, 1:] >= z[0, :-1]).all())                           # the case under test exists
    w = torch.zeros(4, 64)
    w[3] = torch.rand(64)
    merged, zs = ops.sample_pdf_merge(z, w.cuda(), torch.linspace(0.0, 1.0, 128).cuda(), want_samples=True)
    assert torch.equal(merged, torch.sort(torch.cat([z, zs], -1), -1).values)


# ------------------------------------------------------------------------------------------------ rays
def test_rays_against_reference_fixture(golden):
    from nerfpp_b200 import ops
    g = golden("ray_utils.npz")
    h, w = int(g["h"]), int(g["w"])
    ro, rd = ops.get_rays(h, w, g["K"], g["c2w"])
    close(ro.view(h, w, 3), g["rays_o_img"], rtol=0, atol=0)
    close(rd.view(h, w, 3), g["rays_d_img"], rtol=1e-6, atol=1e-7)
    ro2, rd2 = ops.get_rays(h, w, g["K"], g["c2w"], 2, 5)                      # a row tile of the same image
    assert torch.equal(rd2, rd.view(h, w, 3)[2:5].reshape(-1, 3))
    o, d = T(g["o"]).cuda(), T(g["d"]).cuda()
    rb = ops.rays_prepare(o, d, g["bbox"].tolist(), 0.0, True)
    close(rb[:, 6], g["near"], rtol=1e-6, atol=1e-6)
    close(rb[:, 7], g["far"], rtol=1e-6, atol=1e-6)
    ref_rb = O.ray_batch(T(g["o"]), T(g["d"]), T(g["bbox"]))
    close(rb, ref_rb, rtol=1e-6, atol=1e-6)
    # z and points: same floats as the un-fused ATen expressions
    t = torch.linspace(0.0, 1.0, 64)
    z = ops.z_sample(rb, t.cuda())
    zr = rb[:, 6:7].cpu() * (1.0 - t) + rb[:, 7:8].cpu() * t
    assert torch.equal(z.cpu(), zr)
    pts = ops.sample_points(rb, z)
    pr = rb[:, None, 0:3].cpu() + rb[:, None, 3:6].cpu() * z.cpu()[:, :, None]
    assert torch.equal(pts.cpu(), pr)


# ------------------------------------------------------------------------------------------------ encoders
def test_encoders(golden):
    from nerfpp_b200 import ops
    g = golden("encoders.npz")
    x = T(g["x"]).cuda()
    for mr, key in ((10, "emb10"), (4, "emb4")):
        e = ops.posenc(x, O.posenc_freqs(mr))
        close(e, g[key], rtol=1e-5, atol=2e-5)        # sin/cos of arguments up to 2^9 * 1.5: abs error of the argument rounding
        assert torch.equal(e[:, :3], x)
    d = T(g["dirs"]).cuda()
    for deg in (2, 3, 4, 5):
        close(ops.sh_encode(d, deg), g[f"sh{deg}"], rtol=1e-4, atol=2e-6)
    for deg in range(1, 9):                            # degrees 6..8 exist only in the CUDA encoder: closed form
        close(ops.sh_encode(d, deg), O.sh_encode_closed_form(g["dirs"], deg), rtol=1e-4, atol=5e-6)
    # the reference's CUDA kernel outputs (tests/golden/cush.npz), unit and non-unit directions, degrees 1..8
    c = golden("cush.npz")
    for deg in range(1, 9):
        close(ops.sh_encode(T(c["dirs"]).cuda(), deg), c[f"sh{deg}"], rtol=2e-6, atol=1e-6)
    # strided read straight out of a ray batch
    rb = torch.zeros(32, 11, device="cuda")
    rb[:, 8:11] = d
    assert torch.equal(ops.sh_encode(rb[:, 8:11], 4), ops.sh_encode(d, 4))


# ------------------------------------------------------------------------------------------------ loss / optimiser
def test_huber_and_adam():
    from nerfpp_b200 import ops
    torch.manual_seed(2)
    pred, tgt = torch.randn(4096, 3) * 2, torch.rand(4096, 3)
    p = pred.clone().requires_grad_(True)
    ref = O.huber(p, tgt)
    ref.backward()
    loss = torch.zeros(1, device="cuda")
    grad = torch.empty(4096, 3, device="cuda")
    ops.huber_fwd_bwd(pred.cuda(), tgt.cuda(), loss, grad)
    close(loss, ref.detach().reshape(1), rtol=1e-5, atol=1e-7)
    close(grad, p.grad, rtol=1e-6, atol=1e-9)
    ref_t = torch.nn.functional.huber_loss(pred, tgt)
    close(loss, ref_t.reshape(1), rtol=1e-5, atol=1e-7)

    n = 10007
    prm = torch.randn(n)
    m, v = torch.zeros(n), torch.zeros(n)
    pc, gc, mc, vc = prm.cuda(), torch.empty(n, device="cuda"), m.cuda(), v.cuda()
    shadow = torch.empty(n, dtype=torch.float16, device="cuda")
    tp = prm.clone().requires_grad_(True)
    opt = torch.optim.Adam([tp], lr=1e-2, betas=(0.9, 0.99), eps=1e-15)
    p64, m64, v64 = prm.double().numpy(), m.double().numpy(), v.double().numpy()
    for step in range(1, 6):
        g = torch.randn(n) * 1e-3
        g[::7] = 0.0                                   # zero-gradient entries must not move (eps = 1e-15)
        gc.copy_(g)
        ops.adam_step(pc, gc, mc, vc, 1e-2, step, shadow_f16=shadow)
        assert float(gc.abs().max()) == 0.0            # gradient cleared in the same pass
        tp.grad = g.clone()
        opt.step()
        O.adam_step(p64, g.double().numpy(), m64, v64, 1e-2, step)
    close(pc, tp.detach(), rtol=1e-5, atol=1e-6)
    close(pc, p64, rtol=1e-5, atol=1e-6)
    assert torch.equal(shadow, pc.half())
    assert torch.equal(pc[::7].cpu(), prm[::7])


# ------------------------------------------------------------------------------------------------ BASELINE-size properties
def test_full_size_composite_and_sampler_properties():
    """BASELINE C2 sizes (4096 rays, 64 coarse / 192 merged samples): size-independent properties of the rendering kernels —
    weights in [0,1] with sum == acc <= 1, rgb linear in sigmoid(rgb logits) for fixed weights, depth inside [z_min, z_max],
    the backward equals a finite difference of the forward along a random direction, merged z sorted and a superset of the
    coarse z, importance samples inside the coarse span."""
    from nerfpp_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(12)
    R, S, N = 4096, 64, 128
    raw = torch.randn(R, S, 4, generator=g, device="cuda")
    z = 2 + torch.sort(torch.rand(R, S, generator=g, device="cuda") * 4, -1).values
    d = torch.randn(R, 3, generator=g, device="cuda")
    out = ops.composite_fwd(raw, z, d)
    w = out["weights"]
    assert float(w.min()) >= 0.0 and float(w.max()) <= 1.0
    np.testing.assert_allclose(w.sum(-1).cpu().numpy(), out["acc"].cpu().numpy(), rtol=1e-5, atol=1e-6)
    assert float(out["acc"].max()) <= 1.0 + 1e-5
    np.testing.assert_allclose(out["rgb"].cpu().numpy(), (w[..., None] * torch.sigmoid(raw[..., :3])).sum(1).cpu().numpy(), rtol=1e-4, atol=1e-6)
    hit = out["acc"] > 1e-3
    assert bool(((out["depth"] >= z[:, 0] - 1e-4) & (out["depth"] <= z[:, -1] + 1e-4))[hit].all())
    # directional derivative: <d_raw, v> == d/deps sum(rgb * G)(raw + eps v) (central difference in fp32: 1e-2 agreement)
    G = torch.randn(R, 3, generator=g, device="cuda")
    v = torch.randn(R, S, 4, generator=g, device="cuda")
    v[:, -1, 3] = 0.0     # the last interval is 1e10 |d| (NeRFRenderer.h:240): alpha_last jumps 0 -> 1 at sigma = 0, not differentiable there
    d_raw = ops.composite_bwd(raw, z, d, g_rgb=G)
    eps = 2e-3
    fp = (ops.composite_fwd(raw + eps * v, z, d)["rgb"].double() * G).sum()
    fm = (ops.composite_fwd(raw - eps * v, z, d)["rgb"].double() * G).sum()
    lhs, rhs = float((d_raw.double() * v).sum()), float((fp - fm) / (2 * eps))
    assert abs(lhs - rhs) <= 2e-2 * abs(rhs) + 1e-2, (lhs, rhs)     # relu(sigma) kinks inside +-eps|v| add O(eps) noise
    # sampler + merge at full size
    zf, zs = ops.sample_pdf_merge(z, w, torch.linspace(0, 1, N).cuda(), want_samples=True)
    assert zf.shape == (R, S + N) and bool((zf[:, 1:] >= zf[:, :-1]).all())
    assert torch.equal(torch.sort(torch.cat([z, zs], -1), -1).values, zf)           # exactly the reference's sort(cat(...))
    assert bool((zs >= z[:, :1]).all() and (zs <= z[:, -1:]).all())


def test_new_entries_empty_and_ragged():
    """n = 0 and ragged sizes through the fused-point, permutation and one-call render entries."""
    from nerfpp_b200 import ops
    from nerfpp_b200.pipeline import HashNeRF, synthetic_rays
    m = HashNeRF((-1.5, -1.5, -1.5, 1.5, 1.5, 1.5), log2_hashmap_size=14, seed=1)
    empty = torch.empty(0, 3, device="cuda")
    rb0 = ops.rays_prepare(empty, empty, m.bbox, 0.0, True)
    z0 = ops.z_sample(rb0, m.t_vals)
    enc, keep = ops.hash_encode_rays_fwd(m.grid, m.table_f16, rb0, z0)
    assert enc.shape == (0, 32) and keep.shape == (0,)
    ops.hash_encode_rays_bwd(m.grid, rb0, z0, torch.empty(0, 32, device="cuda", dtype=torch.bfloat16), m.grads[:m.n_table])
    out = m.render_rays_fused(empty, empty)
    assert out["rgb"].shape == (0, 3)
    for n in (1, 33, 129):
        o, d, _ = synthetic_rays(n, seed=n)
        a, b = m.render_rays(o, d), m.render_rays_fused(o, d)
        assert torch.equal(a["rgb"], b["rgb"]) and torch.isfinite(a["rgb"]).all()


@pytest.mark.parametrize("r,s,deg,lin", [(1, 1, 4, False), (37, 64, 4, False), (4096, 64, 4, False), (300, 17, 8, True), (129, 33, 1, False)])
def test_ray_setup_equals_the_three_kernels_bit_for_bit(r, s, deg, lin):
    """nrf_ray_setup == nrf_rays_prepare (viewdirs) + nrf_z_sample + nrf_sh_encode_fwd(viewdirs), and resets the scalar it is given."""
    from nerfpp_b200 import ops
    g = torch.Generator().manual_seed(r * 100 + s)
    o = (torch.randn(r, 3, generator=g) * 2.5).cuda()
    d = torch.randn(r, 3, generator=g).cuda()
    bbox = (-1.5, -1.2, -1.0, 1.5, 1.3, 1.1)
    t = torch.linspace(0, 1, s).cuda()
    rb = ops.rays_prepare(o, d, bbox, 0.05, True)
    z = ops.z_sample(rb, t, lin)
    sh = ops.sh_encode(rb[:, 8:11], deg)
    acc = torch.full((1,), 7.0, device="cuda")
    rb2, z2, sh2 = ops.ray_setup(o, d, bbox, 0.05, t, deg, lin_disp=lin, zero_scalar=acc)
    assert torch.equal(rb, rb2) and torch.equal(z, z2) and torch.equal(sh, sh2) and float(acc) == 0.0
    rb3, z3, sh3 = ops.ray_setup(o, d, bbox, 0.05, t, None, lin_disp=lin)
    assert sh3 is None and torch.equal(rb, rb3) and torch.equal(z, z3)


@pytest.mark.parametrize("S,white", [(192, False), (64, True), (33, False), (320, True)])
def test_fused_composite_huber_backward_equals_the_three_entries(S, white):
    """nrf_composite_huber_bwd (the training tail of the colour pass in one launch) against nrf_composite_fwd -> nrf_huber_fwd_bwd ->
    nrf_composite_bwd: same RGB and d_raw bit for bit (the same expressions in the same order), same loss up to the order of the atomic sum."""
    from nerfpp_b200 import ops
    g = torch.Generator().manual_seed(S)
    R = 300
    raw = (torch.randn(R, S, 4, generator=g) * 2).cuda()
    z = (2 + torch.sort(torch.rand(R, S, generator=g) * 4, -1).values).cuda()
    d = torch.randn(R, 3, generator=g).cuda()
    tgt = (torch.rand(R, 3, generator=g) * 3 - 1).cuda()                      # errors beyond delta = 1 included
    out = ops.composite_fwd(raw, z, d, white)
    loss_a, g_rgb = torch.zeros(1, device="cuda"), torch.empty(R, 3, device="cuda")
    ops.huber_fwd_bwd(out["rgb"], tgt, loss_a, g_rgb, 1.0, 0.5)
    d_a = ops.composite_bwd(raw, z, d, white, g_rgb=g_rgb)
    loss_b = torch.zeros(1, device="cuda")
    d_b, rgb_b = ops.composite_huber_bwd(raw, z, d, tgt, loss_b, white_bkgr=white, grad_scale=0.5)
    if S <= 256:
        assert torch.equal(rgb_b, out["rgb"])
        assert torch.equal(d_b, d_a)
    else:       # beyond 256 samples the forward and the two-pass backward are separate looped kernels: same expressions, own summation order
        close(rgb_b, out["rgb"], rtol=1e-5, atol=1e-6)
        close(d_b, d_a, rtol=1e-4, atol=1e-6 * float(d_a.abs().max()))
    assert abs(float(loss_a) - float(loss_b)) <= 1e-6 * abs(float(loss_a))
