"""GPU parity of the fused LeRF language head (nrf_lerf_*, tcgen05) against oracle/restate.py — the restatement of
LeRFImpl::forward (src/LeRF.cpp:28-111), RunLENetwork's keep mask (src/LeRFRenderer.cpp:18-20), RawToLEOutputs (:27-82) and
RenderCLIPEmbedding (src/LeRFRenderer.h:45-54) that tests/test_oracle_pin.py pins against the compiled reference — and against the
fixture generated from the reference itself (tests/golden/lerf.npz).  Floating-point kernel: fp16 operands, fp32 accumulation;
tolerance rel 1e-2 of the output scale (the north star's bf16-class bound).  Inference only in this round."""
import math

import numpy as np
import pytest
import torch

import restate as O

pytestmark = pytest.mark.gpu

TOL = 1e-2
NAMES = ("lang_model_sigma_le_net_0.weight", "lang_model_sigma_le_net_1.weight", "lang_model_le_net_0.weight", "lang_model_le_net_1.weight")
SHAPES = ((256, 128), (33, 256), (256, 160), (512, 256))


def _params(seed=0, gain=1.0):
    g = torch.Generator().manual_seed(seed)
    return {n: (torch.randn(o, i, generator=g) * gain * math.sqrt(2.0 / i)).cuda() for n, (o, i) in zip(NAMES, SHAPES)}


def _enc(n, seed=1):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(n, 128, generator=g).half().cuda()


def _oracle(enc, p, keep=None, dtype=torch.float64):
    w = [p[n].cpu().to(dtype) for n in NAMES]
    out = O.lerf_forward(enc.cpu().to(dtype), w[:2], w[2:])
    return O.lerf_apply_keep(out, keep.cpu().bool()) if keep is not None else out


def _check_raw(out, ref):
    out, ref = out.double().cpu(), ref.double()
    emb_err = float((out[:, :512] - ref[:, :512]).abs().max())
    assert emb_err <= TOL * float(ref[:, :512].abs().max()), emb_err
    cos = (out[:, :512] * ref[:, :512]).sum(-1)
    nz = ref[:, :512].norm(dim=-1) > 0
    assert float(cos[nz].min()) > 1 - 1e-4, float(cos[nz].min())
    sig_err = float((out[:, 512] - ref[:, 512]).abs().max())
    assert sig_err <= TOL * max(float(ref[:, 512].abs().max()), 1e-6), sig_err


def test_forward_matches_reference_fixture(golden):
    """The reference's own LeRF::forward output (fixture from oracle/_ref, fp32 CPU) on its weights and inputs."""
    from nerfpp_b200 import ops
    g = golden("lerf.npz")
    p = {n: torch.from_numpy(g[k]).cuda() for n, k in zip(NAMES, ("sw0", "sw1", "lw0", "lw1"))}
    packed = ops.lerf_pack(p)
    out = ops.lerf_fwd(packed, torch.from_numpy(g["x"]).half().cuda())
    assert out.shape == (96, 513)
    _check_raw(out, torch.from_numpy(g["out"]))
    np.testing.assert_allclose(out[:, :512].norm(dim=-1).cpu().numpy(), 1.0, rtol=1e-5)


@pytest.mark.parametrize("n", [1, 127, 129, 1000, 148 * 128 + 77])
def test_forward_matches_oracle_ragged_and_keep(n):
    from nerfpp_b200 import ops
    p = _params(seed=n)
    enc = _enc(n, seed=n + 1)
    keep = (torch.arange(n) % 5 != 0).to(torch.uint8).cuda()
    packed = ops.lerf_pack(p)
    out = ops.lerf_fwd(packed, enc, keep)
    _check_raw(out, _oracle(enc, p, keep))
    assert float(out[keep == 0, 512].abs().max()) == 0.0
    out2 = ops.lerf_fwd(packed, enc, None)
    assert torch.equal(out2[:, :512], out[:, :512])                       # deterministic; the mask touches sigma only
    assert torch.equal(out2[keep != 0, 512], out[keep != 0, 512])


def test_empty_and_unsupported_shape():
    from nerfpp_b200 import cabi, ops
    p = _params()
    packed = ops.lerf_pack(p)
    assert ops.lerf_fwd(packed, torch.empty(0, 128, dtype=torch.float16, device="cuda")).shape == (0, 513)
    assert ops.lerf_sigma_fwd(packed, torch.empty(0, 128, dtype=torch.float16, device="cuda")).shape == (0, 4)
    with pytest.raises(cabi.NrfError):
        ops.lerf_pack(p, shape=ops.lerf_shape(lang_embed_dim=768))        # src/main.cpp:212 default: not built, must say so


def test_sigma_and_hidden_programs_equal_the_raw_program():
    """The three stage programs share S0 / S1 (and E0): identical MMAs in identical order, so the densities are bit-identical."""
    from nerfpp_b200 import ops
    n = 5000
    p = _params(seed=3)
    enc = _enc(n, seed=4)
    keep = (torch.arange(n) % 7 != 3).to(torch.uint8).cuda()
    packed = ops.lerf_pack(p)
    raw = ops.lerf_fwd(packed, enc, keep)
    raw4 = ops.lerf_sigma_fwd(packed, enc, keep)
    raw4h, hidden, q = ops.lerf_hidden_fwd(packed, enc, keep)
    assert torch.equal(raw4[:, 3], raw[:, 512]) and torch.equal(raw4h, raw4)
    assert float(raw4[:, :3].abs().max()) == 0.0
    # q = |W_e1 h2|^2 through G = W_e1^T W_e1 against the fp64 oracle's un-normalised embedding
    w = [p[k].cpu().double() for k in NAMES]
    x = enc.cpu().double()
    h1 = torch.relu(x @ w[0].t())
    s = h1 @ w[1].t()
    h2 = torch.relu(torch.cat([s[:, 1:], x], -1) @ w[2].t())
    e = h2 @ w[3].t()
    qref = (e * e).sum(-1)
    assert float(((q.cpu().double() - qref).abs() / qref.clamp_min(1e-6)).max()) <= 2 * TOL
    # the h2 tile records: [tile][32 chunks][128 rows][8] fp16
    rec = hidden.view(torch.float16).view(-1, 32, 128, 8).permute(0, 2, 1, 3).reshape(-1, 256)[:n].float().cpu().double()
    assert float((rec - h2).abs().max()) <= TOL * float(h2.abs().max())


def test_large_last_layer_does_not_overflow_the_norm_matrix():
    """G = W_e1^T W_e1 travels in fp16 divided by a power of two chosen by nrf_lerf_pack: with |W_e1| ~ 18 its diagonal (~1.6e5) would overflow
    fp16 unscaled; the norms, the rendered embedding and the raw program stay within tolerance."""
    from nerfpp_b200 import ops
    r, s = 16, 96
    n = r * s
    p = _params(seed=21)
    p[NAMES[3]] = p[NAMES[3]] * 200.0
    p[NAMES[1]] = p[NAMES[1]].clone()
    p[NAMES[1]][0] *= 6.0
    enc = _enc(n, seed=22)
    packed = ops.lerf_pack(p)
    raw4, hidden, q = ops.lerf_hidden_fwd(packed, enc)
    w = [p[k].cpu().double() for k in NAMES]
    x = enc.cpu().double()
    h1 = torch.relu(x @ w[0].t())
    sg = h1 @ w[1].t()
    h2 = torch.relu(torch.cat([sg[:, 1:], x], -1) @ w[2].t())
    e = h2 @ w[3].t()
    qref = (e * e).sum(-1)
    assert float(qref.max()) > 1e6 and bool(torch.isfinite(q).all())
    assert float(((q.cpu().double() - qref).abs() / qref.clamp_min(1e-6)).max()) <= 2 * TOL
    g = torch.Generator().manual_seed(5)
    z = (2 + torch.sort(torch.rand(r, s, generator=g) * 4, -1).values).cuda()
    d = torch.randn(r, 3, generator=g).cuda()
    comp = ops.composite_fwd(raw4.view(r, s, 4), z, d)
    rendered = ops.lerf_render_embedding(packed, comp["weights"], hidden, q)
    ref_raw = _oracle(enc, p, None, torch.float32).view(r, s, 513)
    ref_raw[..., 512] = raw4.view(r, s, 4)[..., 3].cpu()
    ref = O.raw_to_le_outputs(ref_raw, z.cpu(), d.cpu(), 512)["rendered"].double()
    assert float((rendered.cpu().double() - ref).abs().max()) <= TOL * float(ref.abs().max())
    _check_raw(ops.lerf_fwd(packed, enc), _oracle(enc, p))


@pytest.mark.parametrize("r,s", [(64, 192), (7, 64), (3, 200)])
def test_fused_render_matches_raw_to_le_outputs(r, s):
    """Fine pass without the [N,512] embedding: composite(sigma) -> weights, then normalize(W_e1 sum_s (w_s / |e_s|) h2_s) against
    RawToLEOutputs on the oracle's embedding (the oracle is given OUR densities so that the comparison isolates the embedding path from
    sign flips of a near-zero density at the 1e10-wide last interval; the densities themselves are checked above)."""
    from nerfpp_b200 import ops
    n = r * s
    p = _params(seed=r)
    # a density head with both signs and O(1..10) magnitude so that rays terminate inside the interval
    p[NAMES[1]] = p[NAMES[1]].clone()
    p[NAMES[1]][0] *= 6.0
    enc = _enc(n, seed=s)
    keep = (torch.arange(n) % 11 != 0).to(torch.uint8).cuda()
    g = torch.Generator().manual_seed(5)
    z = (2 + torch.sort(torch.rand(r, s, generator=g) * 4, -1).values).cuda()
    d = torch.randn(r, 3, generator=g).cuda()
    packed = ops.lerf_pack(p)
    raw4, hidden, q = ops.lerf_hidden_fwd(packed, enc, keep)
    raw4 = raw4.view(r, s, 4)
    raw4[0, :, 3] = -1.0                                                  # an empty ray: rendered = normalize(0) = 0
    comp = ops.composite_fwd(raw4, z, d)
    rendered = ops.lerf_render_embedding(packed, comp["weights"], hidden, q)
    ref_raw = _oracle(enc, p, keep, torch.float32).view(r, s, 513)
    ref_raw[..., 512] = raw4[..., 3].cpu()
    ref = O.raw_to_le_outputs(ref_raw, z.cpu(), d.cpu(), 512)
    np.testing.assert_allclose(comp["weights"].cpu().numpy(), ref["weights"].numpy(), rtol=1e-3, atol=1e-6)
    np.testing.assert_allclose(comp["depth"].cpu().numpy(), ref["depth"].numpy(), rtol=1e-3, atol=1e-5)
    out, exp = rendered.cpu().double(), ref["rendered"].double()
    assert float(out[0].abs().max()) == 0.0 and float(exp[0].abs().max()) == 0.0
    assert float((out - exp).abs().max()) <= TOL * float(exp.abs().max())
    cos = (out[1:] * exp[1:]).sum(-1)
    assert float(cos.min()) > 1 - 1e-4, float(cos.min())
    # and the same through the compatibility entry + the reference's own formula on OUR raw_le
    raw_le = ops.lerf_fwd(packed, enc, keep).view(r, s, 513).cpu()
    raw_le[0, :, 512] = -1.0
    via_raw = O.raw_to_le_outputs(raw_le, z.cpu(), d.cpu(), 512)["rendered"].double()
    assert float((out - via_raw).abs().max()) <= TOL * float(exp.abs().max())


def test_full_size_properties():
    """BASELINE C5 fine pass: 1024 rays x 192 samples.  Size-independent properties: unit-norm embeddings; positive homogeneity of the
    bias-free ReLU networks (src/LeRF.cpp:12,15) — scaling the input by 2 is exact in fp16 and fp32 except where an activation is an
    fp16 subnormal, so the embedding is unchanged and the density doubles to within that; determinism (bit-exact)."""
    from nerfpp_b200 import ops
    n = 1024 * 192
    p = _params(seed=9)
    enc = _enc(n, seed=10)
    packed = ops.lerf_pack(p)
    out = ops.lerf_fwd(packed, enc)
    norms = out[:, :512].norm(dim=-1)
    assert float((norms - 1).abs().max()) < 1e-4
    out2 = ops.lerf_fwd(packed, (enc.float() * 2).half())
    assert float((out2[:, :512] - out[:, :512]).abs().max()) < 1e-4
    assert float((out2[:, 512] - out[:, 512] * 2).abs().max()) < 1e-3 * float(out[:, 512].abs().max())
    assert torch.equal(ops.lerf_fwd(packed, enc), out)
    idx = torch.randperm(n)[:512]
    _check_raw(out[idx], _oracle(enc[idx], p))


BBOX = (-1.5, -1.5, -1.5, 1.5, 1.5, 1.5)


def _field(seed=3, T=14):
    """A language field with O(1) table entries and He-scaled weights (and a x6 density row) so that the test has signal."""
    from nerfpp_b200.lerf import LeRFField
    f = LeRFField(BBOX, log2_hashmap_size=T, seed=seed)
    g = torch.Generator().manual_seed(seed)
    f.table.copy_((torch.rand(f.n_table, generator=g) * 2 - 1).cuda())
    for k, v in f.weights.items():
        v.copy_((torch.randn(v.shape, generator=g) * math.sqrt(2.0 / v.shape[1])).cuda())
    f.weights["lang_model_sigma_le_net_1.weight"][0] *= 6.0
    f.refresh()
    return f


def _oracle_le_network(f):
    g = f.grid
    g.c_struct()
    meta = dict(box_min=BBOX[:3], box_max=BBOX[3:], scales=g.level_scale.cpu().numpy(), primes=g.primes.cpu().numpy(),
                biases=g.biases.cpu().numpy(), offsets=g.feat_local_idx.cpu().numpy(), sizes=g.feat_local_size.cpu().numpy())
    table = f.table_f16.cpu().numpy()
    w = [f.weights[n].cpu() for n in NAMES]

    def run(pts):                                                            # RunLENetwork, src/LeRFRenderer.cpp:5-25
        r, s, _ = pts.shape
        cl, keep = O.clamp_keep(pts.reshape(-1, 3).numpy(), BBOX[:3], BBOX[3:])
        enc = O.hash_encode(cl, table_f16=table, n_features=8, **meta)
        out = O.lerf_apply_keep(O.lerf_forward(torch.from_numpy(enc), w[:2], w[2:]), torch.from_numpy(keep))
        return out.reshape(r, s, 513)
    return run


def _rays(n, seed=0):
    g = torch.Generator().manual_seed(seed)
    o = torch.tensor([0.3, -0.2, 4.0]).repeat(n, 1) + 0.05 * torch.randn(n, 3, generator=g)
    d = torch.tensor([0.0, 0.0, -1.0]) + 0.25 * torch.randn(n, 3, generator=g)
    return o, d


def test_render_rays_matches_composed_oracle():
    """LeRFRenderer::RenderRays (src/LeRFRenderer.cpp:85-162) through nerfpp_b200/lerf.py against the composed oracle: exact sample counts,
    rendered embedding / depth / acc within the bf16-class tolerance, and the reference's LangEmbedding / Raw on request."""
    f = _field()
    o, d = _rays(24)
    out = f.render_rays(o.cuda(), d.cuda(), return_embedding=True)
    assert out["z"].shape == (24, 64 + 128) and out["rendered"].shape == (24, 512) and out["raw"].shape == (24, 192, 513)
    assert torch.equal(torch.sort(out["z"], -1).values, out["z"])
    rb = O.ray_batch(o, d, torch.tensor(BBOX))
    ref, _, z_ref = O.lerf_render_rays(rb, 64, 128, _oracle_le_network(f), 512)
    zd = (out["z"].cpu() - z_ref).abs()
    print("max |z_fine - oracle| =", float(zd.max()), " median =", float(zd.median()))
    assert float(zd.median()) < 1e-3
    np.testing.assert_allclose(out["depth"].cpu().numpy(), ref["depth"].numpy(), rtol=1e-2, atol=1e-2)
    np.testing.assert_allclose(out["acc"].cpu().numpy(), ref["acc"].numpy(), rtol=1e-2, atol=1e-2)
    hit = ref["acc"] > 1e-3                                                  # rays that miss the box render normalize(0) = 0 on both sides
    cos = (out["rendered"].cpu() * ref["rendered"]).sum(-1)
    print("min cosine(rendered, oracle) =", float(cos[hit].min()))
    assert float(cos[hit].min()) > 1 - 1e-2
    assert float((out["rendered"].cpu() - ref["rendered"]).abs().max()) <= 2e-2
    # the fused rendered embedding equals RenderCLIPEmbedding applied to the compatibility entry's LangEmbedding with the same weights
    via = O.render_clip_embedding(out["embedding"].cpu(), out["weights"].cpu()[..., None])
    assert float((out["rendered"].cpu() - via).abs().max()) <= 1e-2 * float(via.abs().max())


def test_graph_replay_equals_eager():
    f = _field(seed=7)
    o, d = _rays(96, seed=2)
    eager = f.render_rays(o.cuda(), d.cuda(), return_weights=False)
    f.capture_render(96)
    out = f.render_rays_graph(o.pin_memory(), d.pin_memory())
    torch.cuda.synchronize()
    for k in ("rendered", "depth", "acc", "z"):
        assert torch.equal(out[k], eager[k]), k


def test_render_image_tiles_agree():
    """Image rows sharded across ranks (BASELINE C4/C5 rendering): two half-frames equal the whole frame bit for bit."""
    f = _field(seed=5)
    h, w = 16, 24
    K = [[30.0, 0, 12.0], [0, 30.0, 8.0], [0, 0, 1]]
    c2w = [[1, 0, 0, 0.1], [0, 1, 0, -0.1], [0, 0, 1, 4.0]]
    whole = f.render_image(h, w, K, c2w, chunk=100)
    top, bottom = f.render_image(h, w, K, c2w, row_end=7), f.render_image(h, w, K, c2w, row_begin=7)
    for k in ("rendered", "depth", "acc"):
        assert torch.equal(torch.cat([top[k], bottom[k]], 0), whole[k]), k
    assert whole["rendered"].shape == (h * w, 512)
