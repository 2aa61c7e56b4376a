"""Generates the CPU golden fixtures in tests/golden/ by executing the UNMODIFIED reference sources
(oracle/_ref/nerfpp_ref_cpu.so, built by `make -C oracle cpu` from /root/reference/src) on seeded inputs.

    python tests/golden/make_golden.py

The reference ships no tests or known-answer vectors, so these files are the pin for oracle/restate.py
on machines where oracle/_ref is absent.  Each .npz holds inputs AND reference outputs.
"""
import sys
from pathlib import Path

import numpy as np
import torch

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent.parent / "oracle" / "_ref"))
import nerfpp_ref_cpu as R  # noqa: E402

torch.manual_seed(42)  # src/main.cpp:174
np.random.seed(42)
BBOX = torch.tensor([-1.5, -1.5, -1.5, 1.5, 1.5, 1.5])


def npz(name, **kw):
    np.savez_compressed(HERE / name, **{k: (v.detach().cpu().numpy() if isinstance(v, torch.Tensor) else np.asarray(v)) for k, v in kw.items()})
    print("wrote", name)


def rays(n):
    o = torch.tensor([0.3, -0.2, 4.0]).repeat(n, 1) + 0.05 * torch.randn(n, 3)
    d = torch.tensor([0.0, 0.0, -1.0]) + 0.25 * torch.randn(n, 3)
    return o, d


def sample_pdf():
    r, b = 8, 63
    near = torch.rand(r, 1) * 2 + 2
    z = near + torch.linspace(0, 1, 64)[None, :] * (2 + torch.rand(r, 1))
    bins = 0.5 * (z[:, 1:] + z[:, :-1])
    w = torch.rand(r, b - 1) ** 4
    w[0] = 0.0                      # all-zero ray -> uniform pdf via +1e-8
    w[1] = 0.0
    w[1, 17] = 0.9                  # single spike
    w[2, :30] = 0.0                 # flat cdf head (denom < 1e-5 guard)
    out = R.sample_pdf(bins, w, 128, True)
    merged = R.sort_merge(z, out)
    npz("sample_pdf.npz", z=z, bins=bins, weights=w, samples=out, merged=merged)


def raw_to_outputs():
    pipe = R.make_classic(BBOX, 4, 2, 2, 16, True, False)
    r, s = 6, 64
    raw = torch.randn(r, s, 4) * 3
    raw[0, :, 3] = -1.0             # empty ray
    raw[1, 10, 3] = 1e4             # opaque wall (saturation / clamps)
    raw[2, :, 3] = 60.0
    z = 2 + torch.sort(torch.rand(r, s) * 4, -1).values
    o, d = rays(r)
    out = {}
    for white in (False, True):
        x = raw.clone().requires_grad_(True)
        res = pipe.raw_to_outputs(x, z, d, 0.0, white)
        g = {k: torch.randn_like(v) for k, v in res.items()}
        loss = sum((res[k] * g[k]).sum() for k in res)
        loss.backward()
        tag = "w" if white else "b"
        for k, v in res.items():
            out[f"{tag}_{k}"] = v
            out[f"{tag}_g_{k}"] = g[k]
        out[f"{tag}_d_raw"] = x.grad
    npz("raw_to_outputs.npz", raw=raw, z=z, rays_d=d, **out)


def nerf_small():
    pipe = R.make_hash_cpu(BBOX, 2, 2, 4, 2, 4, 4, 2, 64, 15, 3, 64)  # tiny grid; the model is the BASELINE one except input_ch
    # BASELINE input width needs 16 levels x 2 features: build the model through a 16-level embedder instead
    pipe = R.make_hash_cpu(BBOX, 16, 2, 4, 16, 512, 4, 2, 64, 15, 3, 64)
    pipe.init_model()               # Xavier-normal gain 0.1 (Trainable.h:43)
    ws = [p.detach().clone() for p in pipe.model_params()]
    x = torch.cat([torch.rand(64, 32) * 1e-2, torch.randn(64, 16)], -1).requires_grad_(True)
    out = pipe.model(x)
    g = torch.randn_like(out)
    out.backward(g)
    gw = [p.grad.clone() for p in pipe.model_params()]
    # second weight set with O(1) weights so that ReLUs switch and outputs are O(1)
    with torch.no_grad():
        for p in pipe.model_params():
            p.copy_(torch.randn_like(p) * (2.0 / p.shape[1]) ** 0.5)
            p.grad = None
    ws2 = [p.detach().clone() for p in pipe.model_params()]
    x2 = torch.cat([torch.randn(64, 32).half().float(), torch.randn(64, 16)], -1).requires_grad_(True)
    out2 = pipe.model(x2)
    out2.backward(g)
    gw2 = [p.grad.clone() for p in pipe.model_params()]
    kw = {}
    for i in range(5):
        kw[f"w{i}"], kw[f"gw{i}"], kw[f"v{i}"], kw[f"gv{i}"] = ws[i], gw[i], ws2[i], gw2[i]
    npz("nerf_small.npz", x=x.detach(), out=out, g=g, gx=x.grad, x2=x2.detach(), out2=out2, gx2=x2.grad, **kw)


def nerf_classic():
    pipe = R.make_classic(BBOX, 10, 4, 8, 64, True, False)   # 8 layers, skip at 4, width 64 (fixture size), viewdirs
    pipe.init_model()
    names = pipe.model_param_names()
    with torch.no_grad():
        for n, p in zip(names, pipe.model_params()):
            if n.endswith(".bias"):
                p.copy_(0.1 * torch.randn_like(p))
            else:
                p.copy_(torch.randn_like(p) * (2.0 / p.shape[1]) ** 0.5)
    x = torch.randn(16, 63 + 27)
    out = pipe.model(x)
    kw = {n.replace(".", "__"): p for n, p in zip(names, pipe.model_params())}
    npz("nerf_classic.npz", x=x, out=out, **kw)


def encoders():
    x = torch.rand(32, 3) * 3 - 1.5
    d = torch.randn(32, 3)
    d = d / d.norm(dim=-1, keepdim=True)
    npz("encoders.npz", x=x, emb10=R.embedder(x, 10), emb4=R.embedder(x, 4), dirs=d,
        sh2=R.sh_encoder(d, 2), sh3=R.sh_encoder(d, 3), sh4=R.sh_encoder(d, 4), sh5=R.sh_encoder(d, 5))


def ray_utils():
    h, w = 6, 9
    K = torch.tensor([[11.0, 0, 4.5], [0, 10.0, 3.0], [0, 0, 1]])
    ang = 0.3
    c2w = torch.tensor([[np.cos(ang), 0, np.sin(ang), 0.5], [0, 1, 0, -0.25], [-np.sin(ang), 0, np.cos(ang), 4.0]], dtype=torch.float32)
    ro, rd, _ = R.get_rays(h, w, K, c2w)
    o, d = rays(64)
    d[0] = torch.tensor([0.0, 0.0, -1.0])        # axis aligned: exercises 1/(d+1e-6)
    o[1] = torch.tensor([0.0, 0.0, 0.0])         # origin inside the box -> near clamped to 0
    d[2] = torch.tensor([1.0, 0.0, 0.0])         # misses the box -> far = near + 1e-6
    near, far = R.intersect_aabb(o, d, BBOX, 0.0)
    npz("ray_utils.npz", h=h, w=w, K=K, c2w=c2w, rays_o_img=ro.contiguous(), rays_d_img=rd.contiguous(), o=o, d=d, near=near, far=far, bbox=BBOX)


def render_rays_classic():
    """Full RenderRays on CPU through NeRFRenderer<Embedder,Embedder,NeRF> (config C1 shape, width 32 for fixture size)."""
    pipe = R.make_classic(BBOX, 10, 4, 8, 32, True, False)
    names = pipe.model_param_names()
    with torch.no_grad():
        for n, p in zip(names, pipe.model_params()):
            p.copy_(0.1 * torch.randn_like(p) if n.endswith(".bias") else torch.randn_like(p) * (2.0 / p.shape[1]) ** 0.5)
    o, d = rays(12)
    res = pipe.render(o, d, 64, 128, 4096, False, True)
    resw = pipe.render(o, d, 64, 128, 4096, True, True)
    kw = {n.replace(".", "__"): p for n, p in zip(names, pipe.model_params())}
    npz("render_rays_classic.npz", o=o, d=d, bbox=BBOX, rgb=res["rgb"], depth=res["depth"], disp=res["disp"], acc=res["acc"],
        weights=res["weights"], rgb_white=resw["rgb"], **kw)


def trunc_exp():
    x = torch.tensor([-200.0, -100.0, -3.0, 0.0, 1.0, 4.999, 5.0, 7.0, 20.0], requires_grad=True)
    y = R.trunc_exp(x)
    y.backward(torch.ones_like(y))
    npz("trunc_exp.npz", x=x.detach(), y=y, gx=x.grad)


def lerf():
    """LeRF head at the BASELINE C5 shape (src/main.cpp:210-213 with the 512-d language dimension of C5): LeRF(32, 2, 256, 512, 128),
    and RawToLEOutputs on a small ray set, with the autograd gradients of the four weights and of the input."""
    n, d_in, hid, geo, dim = 96, 128, 256, 32, 512
    sw = [torch.randn(hid, d_in) * (2.0 / d_in) ** 0.5, torch.randn(1 + geo, hid) * (2.0 / hid) ** 0.5]
    lw = [torch.randn(hid, geo + d_in) * (2.0 / (geo + d_in)) ** 0.5, torch.randn(dim, hid) * (2.0 / hid) ** 0.5]
    x = torch.randn(n, d_in).half().float()
    out, names = R.lerf_forward(x, sw, lw, geo, hid, dim)
    r, s = 4, 24
    raw = out.reshape(r, s, dim + 1).clone()
    raw[..., -1] = raw[..., -1] * 4
    raw[0, :, -1] = -1.0            # empty ray: the rendered embedding is normalize(0) = 0
    raw[1, 5, -1] = 1e4             # opaque wall
    z = 2 + torch.sort(torch.rand(r, s) * 4, -1).values
    o, d = rays(r)
    res = R.lerf_raw_to_outputs(raw, z, d, dim)
    kw = {f"le_{k}": v for k, v in res.items()}
    npz("lerf.npz", x=x, sw0=sw[0], sw1=sw[1], lw0=lw[0], lw1=lw[1], out=out, names=np.array(names), raw=raw, z=z, rays_d=d, **kw)


def lerf_grads():
    """The language branch of one training iteration on already-encoded samples through the reference's own LeRF + RawToLEOutputs + LibTorch
    autograd (fp64): loss, rendered embedding, d loss / d x in full, and for each weight gradient its first 8 rows + Frobenius norm (size)."""
    g = np.load(HERE / "lerf.npz")
    sw = [torch.from_numpy(g["sw0"]).double(), torch.from_numpy(g["sw1"]).double()]
    lw = [torch.from_numpy(g["lw0"]).double(), torch.from_numpy(g["lw1"]).double()]
    sw[1] = sw[1].clone()
    sw[1][0] *= 4.0
    r, s = 4, 24
    x = torch.from_numpy(g["x"]).double().reshape(r, s, 128)
    z, d = torch.from_numpy(g["z"]).double(), torch.from_numpy(g["rays_d"]).double()
    target = torch.nn.functional.normalize(torch.randn(r, 512, generator=torch.Generator().manual_seed(1), dtype=torch.float64), dim=-1)
    loss, rendered, grads = R.lerf_language_grads(x, sw, lw, z, d, target, 32, 256, 512)
    kw = {}
    for k, gr in zip(("sigma_w0", "sigma_w1", "le_w0", "le_w1"), grads[:4]):
        kw[f"g_{k}_rows"], kw[f"g_{k}_norm"] = gr[:8].clone(), gr.norm()
    npz("lerf_grads.npz", loss=loss, rendered=rendered, target=target, g_x=grads[4], **kw)


if __name__ == "__main__":
    if "--lerf-grads-only" in sys.argv:
        lerf_grads()
        sys.exit(0)
    if "--lerf-only" in sys.argv:      # leaves the other fixtures (and the RNG stream they were drawn from) untouched
        lerf()
        sys.exit(0)
    sample_pdf()
    raw_to_outputs()
    nerf_small()
    nerf_classic()
    encoders()
    ray_utils()
    render_rays_classic()
    trunc_exp()
    lerf()
    lerf_grads()
