"""TEST INFRASTRUCTURE ONLY — CPU restatement of the reference's ray-batch hot path.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module; the product path
(nerfpp_b200/) never does.  Each function follows the reference lines cited in its docstring (paths relative to
/root/reference).  Plain numpy / torch-CPU, dtype-parametric so that float64 can arbitrate tolerances.

Pinning (see tests/test_oracle_pin.py and DESIGN.md §oracle):
  * every ATen stage here is checked against the UNMODIFIED reference code compiled into oracle/_ref/
    (nerfpp_ref_cpu.so) when that is present, and against the committed fixtures in tests/golden/ (generated from it
    by tests/golden/make_golden.py) when it is not;
  * the two CUDA-only reference kernels (CuHashEmbedder, CuSHEncoder) are pinned by fixtures generated on a B200
    from oracle/_ref/nerfpp_ref_cuda.so (tests/golden/make_golden_cuda.py) and, on the GPU box, live.
Third-party arithmetic (cumsum, searchsorted, sort, linspace, sigmoid, exp) is LibTorch's (the reference pins no
version; torch 2.11.0 here) and is called, not restated.
"""
from __future__ import annotations

import math

import numpy as np
import torch

# ------------------------------------------------------------------------------------------------ hash grid


def level_scales(base_resolution: int, finest_resolution: int, n_levels: int) -> np.ndarray:
    """src/CuHashEmbedder.cu:40 in float32.  The reference evaluates this ON THE DEVICE (exp2f/log2f); numpy's float32
    exp2 can differ from the device's by an ulp on the levels that are not exact powers of two, so GPU parity tests
    pass the device-computed scales (fixture tests/golden/level_scales_16_512_16.npy) instead of calling this."""
    f = np.float32
    l2f, l2b = np.log2(f(finest_resolution)), np.log2(f(base_resolution))
    out = np.empty(n_levels, dtype=np.float32)
    for l in range(n_levels):
        out[l] = np.exp2(f(f(f(l2f - l2b) * f(l)) / f(n_levels - 1)) + l2b)
    return out


def hash_cells(points: np.ndarray, box_min, box_max, scales: np.ndarray, primes: np.ndarray, biases: np.ndarray,
               sizes: np.ndarray):
    """src/CuHashEmbedder.cu:44-90: per (point, level) the 8 hashed positions (uint32, bit-exact) and trilinear weights.
    points [N,3] float32 already clamped.  primes int32 [L,1,3].  Returns pos uint32 [N,L,8], w float32 [N,L,8]."""
    f = np.float32
    pts = points.astype(f)
    bmin, bmax = np.asarray(box_min, f), np.asarray(box_max, f)
    q = (pts - bmin) / (bmax - bmin)                      # :44-46 (IEEE division, per axis)
    n, L = pts.shape[0], len(scales)
    pos = np.empty((n, L, 8), dtype=np.uint32)
    w = np.empty((n, L, 8), dtype=f)
    pr = primes.reshape(L, -1, 3)[:, 0, :].astype(np.int64).astype(np.uint32)
    bs = biases.reshape(L, -1, 3)[:, 0, :].astype(f)
    with np.errstate(over="ignore"):
        for l in range(L):
            p = (q * f(scales[l])).astype(f) + bs[l]      # :44-63 (bias is 0 unless RandBias)
            p = p.astype(f)
            fl = np.floor(p)
            ip = fl.astype(np.uint32)                      # :66-68
            a, b, c = (p - fl).astype(f).T                 # :79-81
            k = 0
            for dx in (0, 1):
                for dy in (0, 1):
                    for dz in (0, 1):                      # order 000,001,010,011,100,... (:70-77): z fastest
                        h = ((ip[:, 0] + np.uint32(dx)) * pr[l, 0]) ^ ((ip[:, 1] + np.uint32(dy)) * pr[l, 1]) ^ \
                            ((ip[:, 2] + np.uint32(dz)) * pr[l, 2])
                        pos[:, l, k] = h % np.uint32(sizes[l])
                        wx = a if dx else (f(1) - a)
                        wy = b if dy else (f(1) - b)
                        wz = c if dz else (f(1) - c)
                        w[:, l, k] = (wx * wy).astype(f) * wz   # :83-90
                        k += 1
    return pos, w


def hash_encode(points, box_min, box_max, scales, primes, biases, offsets, sizes, table_f16: np.ndarray, n_features: int):
    """src/CuHashEmbedder.cu:93-101 + :253,274: fp32 interpolation of fp16 features, rounded to fp16, returned as fp32.
    `offsets` are SCALAR offsets into the flattened table (:55, feat_pool + feat_local_idx[level]) — the reference quirk
    that makes consecutive levels overlap."""
    pos, w = hash_cells(points, box_min, box_max, scales, primes, biases, sizes)
    n, L, _ = pos.shape
    flat = table_f16.reshape(-1)
    out = np.zeros((n, L * n_features), dtype=np.float32)
    for l in range(L):
        for k in range(n_features):
            idx = np.int64(offsets[l]) + pos[:, l, :].astype(np.int64) * n_features + k
            feat = flat[idx].astype(np.float32)            # [N,8]
            acc = np.zeros(n, dtype=np.float32)
            for d in range(8):
                acc = acc + w[:, l, d] * feat[:, d]
            out[:, l * n_features + k] = acc.astype(np.float16).astype(np.float32)
    return out


def hash_encode_unrounded_f64(points, box_min, box_max, scales, primes, biases, offsets, sizes, table_f16, n_features):
    """float64 interpolation without the fp16 output rounding (tolerance arbitration)."""
    pos, w = hash_cells(points, box_min, box_max, scales, primes, biases, sizes)
    n, L, _ = pos.shape
    flat = table_f16.reshape(-1).astype(np.float64)
    out = np.zeros((n, L * n_features))
    for l in range(L):
        for k in range(n_features):
            idx = np.int64(offsets[l]) + pos[:, l, :].astype(np.int64) * n_features + k
            out[:, l * n_features + k] = (w[:, l, :].astype(np.float64) * flat[idx]).sum(-1)
    return out


def hash_encode_bwd_f64(points, box_min, box_max, scales, primes, biases, offsets, sizes, grad_enc: np.ndarray,
                        n_features: int, table_scalars: int) -> np.ndarray:
    """Exact (float64) adjoint of the interpolation: what src/CuHashEmbedder.cu:188-201 accumulates with x128-scaled
    fp16 atomics.  Returns the flat [table_scalars] gradient."""
    pos, w = hash_cells(points, box_min, box_max, scales, primes, biases, sizes)
    n, L, _ = pos.shape
    g = np.zeros(table_scalars, dtype=np.float64)
    for l in range(L):
        for k in range(n_features):
            idx = np.int64(offsets[l]) + pos[:, l, :].astype(np.int64) * n_features + k
            np.add.at(g, idx.reshape(-1), (w[:, l, :].astype(np.float64) * grad_enc[:, l * n_features + k, None].astype(np.float64)).reshape(-1))
    return g


def clamp_keep(points: np.ndarray, box_min, box_max):
    """src/CuHashEmbedder.cpp:92-94,101."""
    bmin, bmax = np.asarray(box_min, np.float32), np.asarray(box_max, np.float32)
    c = np.maximum(np.minimum(points, bmax), bmin)
    keep = (points == c).all(-1)
    return c.astype(np.float32), keep


# ------------------------------------------------------------------------------------------------ encoders


def sh_encode_closed_form(dirs: np.ndarray, degree: int) -> np.ndarray:
    """Real spherical harmonics up to `degree` (exclusive band count, out dim degree^2) from the textbook definition
    Y_lm = (-1)^m sqrt(2) K_l^|m| {cos(m phi) | sin(|m| phi)} P_l^|m|(cos theta)  (P without Condon-Shortley phase),
    index l^2+l+m, evaluated in float64.  For UNIT vectors this equals the polynomial table of
    src/CuSHEncoder.cu:26-103 (and src/NeRF.cpp:155-194 up to degree 5) — an independent derivation of it."""
    d = dirs.astype(np.float64)
    x, y, z = d[:, 0], d[:, 1], d[:, 2]
    n = d.shape[0]
    out = np.zeros((n, degree * degree))
    r_xy = np.sqrt(x * x + y * y)
    phi = np.arctan2(y, x)
    ct = z
    # associated Legendre without Condon-Shortley phase
    P = {}
    for m in range(degree):
        pmm = np.ones(n)
        for i in range(1, m + 1):
            pmm = pmm * (2 * i - 1) * r_xy
        P[(m, m)] = pmm
        if m + 1 < degree:
            P[(m + 1, m)] = ct * (2 * m + 1) * pmm
        for l in range(m + 2, degree):
            P[(l, m)] = ((2 * l - 1) * ct * P[(l - 1, m)] - (l + m - 1) * P[(l - 2, m)]) / (l - m)
    for l in range(degree):
        for m in range(-l, l + 1):
            am = abs(m)
            K = math.sqrt((2 * l + 1) / (4 * math.pi) * math.factorial(l - am) / math.factorial(l + am))
            if m == 0:
                v = K * P[(l, 0)]
            elif m > 0:
                v = (-1) ** m * math.sqrt(2) * K * np.cos(m * phi) * P[(l, am)]
            else:
                v = (-1) ** m * math.sqrt(2) * K * np.sin(am * phi) * P[(l, am)]
            out[:, l * l + l + m] = v
    return out


def posenc(x: torch.Tensor, num_freqs: int, max_freq_log2: float | None = None) -> torch.Tensor:
    """src/NeRF.cpp:4-39: [x, sin(f0 x), cos(f0 x), ...], f_k = powf(2, max_freq/(n-1)*k) (log sampling)."""
    max_freq = float(num_freqs - 1) if max_freq_log2 is None else max_freq_log2
    out = [x]
    for f in posenc_freqs(num_freqs, max_freq):
        out.append(torch.sin(x * f))
        out.append(torch.cos(x * f))
    return torch.cat(out, -1)


def posenc_freqs(num_freqs: int, max_freq_log2: float | None = None):
    max_freq = float(num_freqs - 1) if max_freq_log2 is None else max_freq_log2
    f32 = np.float32
    return [float(np.power(f32(2.0), f32(f32(max_freq) / f32(num_freqs - 1)) * f32(i))) for i in range(num_freqs)]


# ------------------------------------------------------------------------------------------------ models


def nerf_small_forward(x: torch.Tensor, weights, input_ch: int = 32, input_ch_views: int = 16) -> torch.Tensor:
    """src/NeRF.cpp:363-412 (no normals): weights = [sigma_net_0, sigma_net_1, ..., color_net_0, ...] as a pair of lists
    (sigma_ws, color_ws), each torch Linear.weight [out, in], bias-free.  Output [rgb(3), sigma]."""
    sigma_ws, color_ws = weights
    pts, views = x[..., :input_ch], x[..., input_ch:input_ch + input_ch_views]
    h = pts
    for i, w in enumerate(sigma_ws):
        h = h @ w.t()
        if i != len(sigma_ws) - 1:
            h = torch.relu(h)
    sigma, geo = h[..., 0], h[..., 1:]
    h = torch.cat([views, geo], -1)                        # views first (:383)
    for i, w in enumerate(color_ws):
        h = h @ w.t()
        if i != len(color_ws) - 1:
            h = torch.relu(h)
    return torch.cat([h, sigma.unsqueeze(-1)], -1)


def nerf_forward(x: torch.Tensor, p: dict, input_ch: int = 63, input_ch_views: int = 27, skips=(4,)) -> torch.Tensor:
    """src/NeRF.cpp:92-126 with use_viewdirs: p maps the reference's registered names (model_pts_linears_i etc.,
    src/NeRF.cpp:76-89) + '.weight'/'.bias' to tensors."""
    pts, views = x[..., :input_ch], x[..., input_ch:input_ch + input_ch_views]
    h = pts
    i = 0
    while f"model_pts_linears_{i}.weight" in p:
        h = torch.relu(h @ p[f"model_pts_linears_{i}.weight"].t() + p[f"model_pts_linears_{i}.bias"])
        if i in skips:
            h = torch.cat([pts, h], -1)
        i += 1
    alpha = h @ p["model_alpha_linear.weight"].t() + p["model_alpha_linear.bias"]
    feat = h @ p["model_feature_linear.weight"].t() + p["model_feature_linear.bias"]
    h = torch.cat([feat, views], -1)
    h = torch.relu(h @ p["model_views_linears_0.weight"].t() + p["model_views_linears_0.bias"])
    rgb = h @ p["model_rgb_linear.weight"].t() + p["model_rgb_linear.bias"]
    return torch.cat([rgb, alpha], -1)


# ------------------------------------------------------------------------------------------------ rendering


class TruncExp(torch.autograd.Function):
    """src/CustomOps.cpp:5-16: forward exp(x) (NOT truncated), backward g * exp(clamp(x, -100, 5))."""

    @staticmethod
    def forward(ctx, x):
        ctx.save_for_backward(x)
        return torch.exp(x)

    @staticmethod
    def backward(ctx, g):
        (x,) = ctx.saved_tensors
        return g * torch.exp(x.clamp(-100.0, 5.0))


def raw_to_outputs(raw: torch.Tensor, z: torch.Tensor, rays_d: torch.Tensor, raw_noise_std: float = 0.0,
                   white_bkgr: bool = False, noise: torch.Tensor | None = None) -> dict:
    """src/NeRFRenderer.h:199-282."""
    dists = z[..., 1:] - z[..., :-1]                                                       # :239
    dists = torch.cat([dists, torch.full_like(dists[..., :1], 1e10)], -1)                   # :240
    dists = dists * torch.norm(rays_d[..., None, :], 2, -1)                                 # :241
    rgb = torch.sigmoid(raw[..., :3])                                                       # :250
    dens = raw[..., 3]                                                                      # :252
    if raw_noise_std > 0.0:
        dens = dens + (noise if noise is not None else torch.randn_like(dens)) * raw_noise_std   # :253-254
    alpha = -TruncExp.apply(-torch.relu(dens) * dists) + 1.0                                # :234,256
    logt = torch.cat([torch.zeros_like(alpha[:, :1]), torch.cumsum(torch.log(torch.clamp_min(1.0 - alpha, 1e-10)), -1)], -1)[:, :-1]   # :263-266
    weights = alpha * TruncExp.apply(logt)                                                  # :267
    rgb_map = torch.sum(weights[..., None] * rgb, -2)                                       # :271
    depth = torch.sum(weights * z, -1) / torch.clamp_min(torch.sum(weights, -1), 1e-10)     # :272
    disp = 1.0 / torch.max(1e-10 * torch.ones_like(depth), depth)                           # :273
    acc = torch.sum(weights, -1)                                                            # :274
    if white_bkgr:
        rgb_map = rgb_map + (1.0 - acc[..., None])                                          # :276-277
    return {"rgb": rgb_map, "depth": depth, "disp": disp, "acc": acc, "weights": weights}


def lerf_forward(x: torch.Tensor, sigma_w, le_w) -> torch.Tensor:
    """src/LeRF.cpp:78-110 (the live branch: an independent density net for the language field, then the language net on
    cat(geo_feat_le, inputs_le)).  sigma_w / le_w: lists of torch Linear weights [out, in]; every layer is bias-free
    (src/LeRF.cpp:12,15).  Returns [N, lang_embed_dim + 1] = [normalised embedding, sigma_le]."""
    h = x
    for i, w in enumerate(sigma_w):                                                         # :86-91
        h = h @ w.t()
        if i != len(sigma_w) - 1:
            h = torch.relu(h)
    sigma_le = h[..., 0]                                                                    # :92
    geo = h[..., 1:]                                                                        # :93
    h = torch.cat([geo, x], -1)                                                             # :96
    for i, w in enumerate(le_w):                                                            # :97-102
        h = h @ w.t()
        if i != len(le_w) - 1:
            h = torch.relu(h)
    le = torch.nn.functional.normalize(h, dim=-1, eps=1e-8)                                 # :105
    return torch.cat([le, sigma_le[..., None]], -1)                                         # :107-110


def lerf_apply_keep(raw_le: torch.Tensor, keep: torch.Tensor) -> torch.Tensor:
    """src/LeRFRenderer.cpp:18-20: sigma_le (last channel) := 0 where the point was outside the box."""
    out = raw_le.clone()
    out[~keep, -1] = 0
    return out


def render_clip_embedding(embeds: torch.Tensor, weights: torch.Tensor) -> torch.Tensor:
    """src/LeRFRenderer.h:45-54.  embeds [R,S,D], weights [R,S,1]."""
    return torch.nn.functional.normalize(torch.sum(weights * embeds, -2), dim=-1, eps=1e-8)


def raw_to_le_outputs(raw_le: torch.Tensor, z: torch.Tensor, rays_d: torch.Tensor, lang_embed_dim: int) -> dict:
    """src/LeRFRenderer.cpp:27-82 with raw_noise_std = 0 (Relevancy, :79, is RuCLIP's and not restated)."""
    dists = z[..., 1:] - z[..., :-1]                                                        # :39
    dists = torch.cat([dists, torch.full_like(dists[..., :1], 1e10)], -1)                   # :40
    dists = dists * torch.norm(rays_d[..., None, :], 2, -1)                                 # :41
    emb = raw_le[..., :lang_embed_dim]                                                      # :46
    dens = raw_le[..., lang_embed_dim]                                                      # :49
    alpha = -TruncExp.apply(-torch.relu(dens) * dists) + 1.0                                # :37,53
    logt = torch.cat([torch.zeros_like(alpha[:, :1]), torch.cumsum(torch.log(torch.clamp_min(1.0 - alpha, 1e-10)), -1)], -1)[:, :-1]   # :62-65
    weights = alpha * TruncExp.apply(logt)                                                  # :66
    depth = torch.sum(weights * z, -1) / torch.clamp_min(torch.sum(weights, -1), 1e-10)     # :70
    disp = 1.0 / torch.max(1e-10 * torch.ones_like(depth), depth)                           # :71
    acc = torch.sum(weights, -1)                                                            # :72
    rendered = render_clip_embedding(emb, weights[..., None])                               # :75
    return {"rendered": rendered, "weights": weights, "depth": depth, "disp": disp, "acc": acc}


def sample_pdf(bins: torch.Tensor, weights: torch.Tensor, n_samples: int, det: bool = True, u: torch.Tensor | None = None,
               sums: str = "torch"):
    """src/Sampler.h:6-43.  Returns (samples, inds) — inds are the searchsorted indices (int64) for exactness checks.

    sums: the summation ORDER of torch::sum / torch::cumsum (:11-12) is implementation-defined in LibTorch (CPU: vectorised
    cascade fp32 sum, cumsum accumulating in double; CUDA: blocked fp32 reduce / scan).  "torch" uses this process's
    torch (what the compiled reference does on CPU); "exact" fixes the order-dependence by making every sum the correctly
    rounded one (fp64 accumulate, one rounding to fp32, knots made monotone) — the variant the CUDA kernel implements.
    The two differ only in the last ulp of a knot, i.e. in `inds` only where u ties a knot within that ulp."""
    weights = weights + 1e-8                                                                # :10
    if sums == "exact":
        total = weights.double().sum(-1, True).to(weights.dtype)
        pdf = weights / total
        cdf = torch.cummax(torch.cumsum(pdf.double(), -1).to(weights.dtype), -1).values
    else:
        pdf = weights / torch.sum(weights, -1, True)                                        # :11
        cdf = torch.cumsum(pdf, -1)                                                         # :12
    cdf = torch.cat([torch.zeros_like(cdf[..., :1]), cdf], -1)                              # :13
    if u is None:
        assert det
        u = torch.linspace(0.0, 1.0, n_samples, dtype=torch.float32).to(cdf.dtype)           # :20
    u = u.expand(list(cdf.shape[:-1]) + [n_samples]).contiguous()                           # :21,27
    inds = torch.searchsorted(cdf, u, right=True)                                           # :28
    below = torch.clamp_min(inds - 1, 0)                                                    # :29
    above = torch.clamp_max(inds, cdf.shape[-1] - 1)                                        # :30
    cdf_g0, cdf_g1 = torch.gather(cdf, -1, below), torch.gather(cdf, -1, above)             # :33-34
    bins_g0, bins_g1 = torch.gather(bins, -1, below), torch.gather(bins, -1, above)         # :35
    denom = cdf_g1 - cdf_g0                                                                 # :37
    denom = torch.where(denom < 1e-5, torch.ones_like(denom), denom)                        # :38
    t = (u - cdf_g0) / denom                                                                # :39
    return bins_g0 + t * (bins_g1 - bins_g0), inds                                          # :40


def get_rays(h: int, w: int, K: torch.Tensor, c2w: torch.Tensor):
    """src/RayUtils.h:5-46 (cone angle omitted: unused with ThinRay)."""
    ys = torch.linspace(0, h - 1, h).view(h, 1).expand(h, w)
    xs = torch.linspace(0, w - 1, w).view(1, w).expand(h, w)
    dirs = torch.stack([(xs - K[0, 2]) / K[0, 0], -(ys - K[1, 2]) / K[1, 1], -torch.ones_like(xs)], -1)
    rays_d = torch.sum(dirs[..., None, :] * c2w[:3, :3], -1)
    rays_o = c2w[:3, -1].expand(rays_d.shape)
    return rays_o, rays_d


def get_ray_batch(rand_h: torch.Tensor, rand_w: torch.Tensor, K: torch.Tensor, c2w: torch.Tensor):
    """src/NeRFDataset.cpp:109-144 (NeRFDataset::GetRayBatch): rays of a list of pixel coordinates, and the scalar cone angle."""
    fx, fy, cx, cy = K[0, 0], K[1, 1], K[0, 2], K[1, 2]
    dirsx = (rand_w.to(torch.float32) - cx) / fx                                             # :122
    dirsy = -(rand_h.to(torch.float32) - cy) / fy                                            # :123
    dirsz = -torch.ones_like(dirsx)                                                          # :124
    dirs = torch.stack([dirsx, dirsy, dirsz], -1)                                            # :126
    rays_d = torch.sum(dirs[..., None, :] * c2w[:3, :3], -1)                                 # :128-131
    rays_o = c2w[:3, -1].expand(rays_d.shape)                                                # :133
    cone_angle = (1.0 / fx + 1.0 / fy) / 2.0                                                 # :137-141
    return rays_o, rays_d, cone_angle


def intersect_aabb(rays_o: torch.Tensor, rays_d: torch.Tensor, bbox: torch.Tensor, near_plane: float = 0.0):
    """src/RayUtils.h:87-126."""
    aabb = bbox.reshape(2, 3)
    frac = 1.0 / (rays_d + 1e-6)
    t_lo = (aabb[0] - rays_o) * frac
    t_hi = (aabb[1] - rays_o) * frac
    nears = torch.max(torch.minimum(t_lo, t_hi), -1).values
    fars = torch.min(torch.maximum(t_lo, t_hi), -1).values
    nears = torch.clamp_min(nears, near_plane)
    fars = torch.maximum(fars, nears + 1e-6)
    return nears, fars


def ray_batch(rays_o, rays_d, bbox, use_viewdirs=True):
    """src/NeRFRenderer.h:549-583: [o, d, near, far (, viewdirs)]."""
    near, far = intersect_aabb(rays_o, rays_d, bbox)
    parts = [rays_o, rays_d, near[:, None], far[:, None]]
    if use_viewdirs:
        parts.append(rays_d / torch.norm(rays_d, 2, -1, True))
    return torch.cat(parts, -1)


def z_coarse(ray_batch_: torch.Tensor, n_samples: int) -> torch.Tensor:
    """src/NeRFRenderer.h:393-397."""
    t = torch.linspace(0.0, 1.0, n_samples, dtype=torch.float32).to(ray_batch_.dtype)
    near, far = ray_batch_[:, 6:7], ray_batch_[:, 7:8]
    return near * (1.0 - t) + far * t


def render_rays(ray_batch_: torch.Tensor, n_samples: int, n_importance: int, run_network, white_bkgr: bool = False):
    """src/NeRFRenderer.h:366-459 in the parity configuration (ThinRay, perturb 0, no noise):
    run_network(pts [R,S,3], viewdirs [R,3]) -> raw [R,S,4].  Returns (fine outputs, coarse outputs, z_fine)."""
    o, d, vd = ray_batch_[:, 0:3], ray_batch_[:, 3:6], ray_batch_[:, 8:11]
    z = z_coarse(ray_batch_, n_samples)
    pts = o[:, None, :] + d[:, None, :] * z[:, :, None]                                     # :419
    out1 = raw_to_outputs(run_network(pts, vd), z, d, 0.0, white_bkgr)                      # :422-423
    z_mid = 0.5 * (z[:, 1:] + z[:, :-1])                                                    # :427
    z_s, _ = sample_pdf(z_mid, out1["weights"][:, 1:-1], n_importance, True)                # :428
    z_all, _ = torch.sort(torch.cat([z, z_s.detach()], -1), -1)                             # :429-431
    pts = o[:, None, :] + d[:, None, :] * z_all[:, :, None]                                 # :432
    out2 = raw_to_outputs(run_network(pts, vd), z_all, d, 0.0, white_bkgr)                  # :447-448
    return out2, out1, z_all


def lerf_language_loss(rendered: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
    """src/NeRFExecutor.h:964-968: huber(delta 1.25, no reduction).sum(-1).nanmean()."""
    return torch.nn.functional.huber_loss(rendered, target, reduction="none", delta=1.25).sum(-1).nanmean()


def lerf_backward_fused_form(x: torch.Tensor, sigma_w, le_w, z: torch.Tensor, rays_d: torch.Tensor, target: torch.Tensor, emulate: bool = False) -> dict:
    """The gradients of the language loss written the way the fused kernels evaluate them: the [N,512] embedding is
    never formed; everything per sample is 256-wide.  x [R,S,C] (the hash encoding of the fine pass), 2-layer nets (BASELINE C5).
    Explicit formulas, no autograd except for the compositing weights (an existing kernel, nrf_composite_bwd).  Returns d loss / d
    {sigma_w0, sigma_w1, le_w0, le_w1, x}; tests compare it with autograd through lerf_forward + raw_to_le_outputs.
    emulate=True additionally applies the rounding points of the sm_100a kernels (csrc/lerf_tc.cu TRAIN program, csrc/lerf_bwd_tc.cu), so that the
    ReLU active sets are the kernels' own: the forward runs on fp16 operands (W_s0, W_s1, W_e0, G / 2^k and h1, geo, h2), the gradient chain
    on bf16 operands (W_e0, W_s1, W_s0 and the gradient rows d a2, d s, d a1, beta h2, d x), the weight-gradient products on bf16 copies of the
    activations ([geo | x], h1, h2 rounded from the fp32 accumulators); W_e1 enters in fp32.  Pass the UNROUNDED weights; x must hold
    fp16-representable values (it is the fp16 output of the hash encoder)."""
    rh = (lambda t: t.float().half().to(t.dtype)) if emulate else (lambda t: t)
    rb = (lambda t: t.float().bfloat16().to(t.dtype)) if emulate else (lambda t: t)
    r_, s_, c_ = x.shape
    xf = x.reshape(-1, c_)
    w_s0, w_s1 = sigma_w
    w_e0, w_e1 = le_w
    a1 = xf @ rh(w_s0).t()
    h1 = rh(torch.relu(a1))
    sg = h1 @ rh(w_s1).t()                                        # [sigma | geo]
    geo = rh(sg[:, 1:])
    gx = torch.cat([geo, xf], -1)
    a2 = gx @ rh(w_e0).t()
    h2 = rh(torch.relu(a2))
    # bf16 copies for the weight-gradient products (identical to the values above when not emulating)
    h1_b, gx_b, h2_b = rb(torch.relu(a1)), torch.cat([rb(sg[:, 1:]), rb(xf)], -1), rb(torch.relu(a2))
    gram = w_e1.t() @ w_e1                                        # G
    if emulate:                                                   # the operand travels as fp16(G / 2^k), 2^k >= max diag / 256 (lerf_gscale_kernel)
        gs = 1.0
        while float(gram.diagonal().max()) > 256.0 * gs:
            gs *= 2.0
        gram = rh(gram / gs) * gs
    t = h2 @ gram                                                 # G h2 (symmetric)
    n2 = (t * h2).sum(-1)                                         # |e|^2 = h2^T G h2
    n = n2.clamp_min(0).sqrt().clamp_min(1e-8)
    sigma = sg[:, 0].reshape(r_, s_).detach().requires_grad_(True)
    raw4 = torch.cat([torch.zeros(r_, s_, 3, dtype=x.dtype), sigma[..., None]], -1)
    w = raw_to_outputs(raw4, z, rays_d)["weights"]                # the compositing weights of the density column
    c = (w.detach() / n.reshape(r_, s_))
    hs = (c[..., None] * h2.reshape(r_, s_, -1)).sum(1)           # [R,256]
    e_ray = hs @ w_e1.t()                                         # E = W_e1 Hs
    e_norm = e_ray.norm(dim=-1, keepdim=True).clamp_min(1e-8)
    rendered = e_ray / e_norm
    # loss -> d rendered (huber, delta 1.25, summed over channels, mean over rays)
    diff = rendered - target
    d_r = torch.where(diff.abs() <= 1.25, diff, 1.25 * torch.sign(diff)) / r_
    d_e = (d_r - rendered * (rendered * d_r).sum(-1, keepdim=True)) / e_norm                        # normalize backward, per ray
    u = d_e @ w_e1                                                                                  # W_e1^T dE   [R,256]
    u_rows = u[:, None, :].expand(r_, s_, -1).reshape(-1, u.shape[-1])
    d_w = ((h2 * u_rows).sum(-1) / n).reshape(r_, s_)                                               # d loss / d w_s
    (d_sigma,) = torch.autograd.grad(w, sigma, d_w)                                                 # compositing backward (existing kernel)
    cf, dwf = c.reshape(-1), d_w.reshape(-1)
    beta = cf * dwf / n                                                                              # W_e1^T e_hat = G h2 / n
    d_h2 = cf[:, None] * u_rows - beta[:, None] * t
    d_we1 = d_e.t() @ hs - w_e1 @ (rb(beta[:, None] * h2).t() @ h2_b)                               # outer term - W_e1 * weighted Gram
    d_a2 = rb(d_h2 * ((h2 > 0) if emulate else (a2 > 0)))
    d_we0 = d_a2.t() @ gx_b
    d_gx = d_a2 @ rb(w_e0)
    d_s = rb(torch.cat([d_sigma.reshape(-1, 1), d_gx[:, :sg.shape[1] - 1]], -1))
    d_ws1 = d_s.t() @ h1_b
    d_a1 = rb((d_s @ rb(w_s1)) * ((h1 > 0) if emulate else (a1 > 0)))
    d_ws0 = d_a1.t() @ gx_b[:, sg.shape[1] - 1:]
    d_x = rb(d_a1 @ rb(w_s0) + d_gx[:, sg.shape[1] - 1:])
    return {"sigma_w0": d_ws0, "sigma_w1": d_ws1, "le_w0": d_we0, "le_w1": d_we1, "x": d_x.reshape(r_, s_, c_), "rendered": rendered}


def lerf_render_rays(ray_batch_: torch.Tensor, n_samples: int, n_importance: int, run_le_network, lang_embed_dim: int):
    """src/LeRFRenderer.cpp:85-162 in the parity configuration (ThinRay, perturb 0, no noise, no preconditioning):
    run_le_network(pts [R,S,3]) -> raw_le [R,S,D+1].  Returns (fine outputs, coarse outputs, z_fine)."""
    o, d = ray_batch_[:, 0:3], ray_batch_[:, 3:6]
    z = z_coarse(ray_batch_, n_samples)                                                     # :112-118
    pts = o[:, None, :] + d[:, None, :] * z[:, :, None]                                     # :137
    out1 = raw_to_le_outputs(run_le_network(pts), z, d, lang_embed_dim)                     # :140-141
    z_mid = 0.5 * (z[:, 1:] + z[:, :-1])                                                    # :145
    z_s, _ = sample_pdf(z_mid, out1["weights"][:, 1:-1], n_importance, True)                # :146
    z_all, _ = torch.sort(torch.cat([z, z_s.detach()], -1), -1)                             # :147-149
    pts = o[:, None, :] + d[:, None, :] * z_all[:, :, None]                                 # :150
    out2 = raw_to_le_outputs(run_le_network(pts), z_all, d, lang_embed_dim)                 # :166-167
    return out2, out1, z_all


# ------------------------------------------------------------------------------------------------ training glue


def huber(pred: torch.Tensor, target: torch.Tensor, delta: float = 1.0) -> torch.Tensor:
    """torch::nn::functional::huber_loss defaults (src/NeRFExecutor.h:883-886): mean reduction, delta 1."""
    e = pred - target
    ae = e.abs()
    return torch.where(ae < delta, 0.5 * e * e, delta * (ae - 0.5 * delta)).mean()


def adam_step(p: np.ndarray, g: np.ndarray, m: np.ndarray, v: np.ndarray, lr: float, step: int, beta1=0.9, beta2=0.99, eps=1e-15):
    """torch::optim::Adam as configured at src/NeRFExecutor.h:539 (no amsgrad / weight decay), float64 maths."""
    m[:] = beta1 * m + (1 - beta1) * g
    v[:] = beta2 * v + (1 - beta2) * g * g
    bc1, bc2 = 1 - beta1 ** step, 1 - beta2 ** step
    denom = np.sqrt(v) / math.sqrt(bc2) + eps
    p[:] = p - (lr / bc1) * m / denom
    return p


def nerf_train_reference(x, p, gout, emulate):
    """fp64 autograd of NeRFImpl::forward (src/NeRF.cpp:92-126, the graph of nerf_forward above with the intermediates kept).
    emulate=True rounds weights, inputs and every stored activation to bf16 (straight-through) — the rounding points of
    nrf_mlp_nerf_fwd_train / nrf_mlp_nerf_bwd — so the ReLU masks are the kernels' masks.  Returns out, grads (by parameter name) and
    `bound` = max over entries of sum |terms| of each gradient (the scale its rounding error is relative to)."""
    rb = (lambda t: t + (t.bfloat16().double() - t).detach()) if emulate else (lambda t: t)
    pd = {k: v.double().requires_grad_(True) for k, v in p.items()}
    q = {k: (rb(v) if k.endswith("weight") else v) for k, v in pd.items()}
    xd = rb(x.double())
    pts, views = xd[:, :63], xd[:, 63:]
    h, hs, pre, xin = pts, {}, {}, {}
    for i in range(8):
        xin[f"model_pts_linears_{i}"] = h
        pre[i] = h @ q[f"model_pts_linears_{i}.weight"].t() + q[f"model_pts_linears_{i}.bias"]          # src/NeRF.cpp:100
        pre[i].retain_grad()
        h = rb(torch.relu(pre[i]))
        hs[i + 1] = h
        if i == 4:
            h = torch.cat([pts, h], -1)                                                               # :103-104
    xin["model_alpha_linear"] = xin["model_feature_linear"] = h
    alpha = h @ q["model_alpha_linear.weight"].t() + q["model_alpha_linear.bias"]                       # :110
    alpha.retain_grad()
    feat = rb(h @ q["model_feature_linear.weight"].t() + q["model_feature_linear.bias"])                # :111
    feat.retain_grad()
    xin["model_views_linears_0"] = torch.cat([feat, views], -1)
    prev = xin["model_views_linears_0"] @ q["model_views_linears_0.weight"].t() + q["model_views_linears_0.bias"]   # :112-116
    prev.retain_grad()
    hv = rb(torch.relu(prev))
    xin["model_rgb_linear"] = hv
    rgb = hv @ q["model_rgb_linear.weight"].t() + q["model_rgb_linear.bias"]                            # :119
    rgb.retain_grad()
    out = torch.cat([rgb, alpha], -1)                                                                   # :120
    (out * gout.double()).sum().backward()
    dy = {f"model_pts_linears_{i}": pre[i].grad for i in range(8)}
    dy.update({"model_alpha_linear": alpha.grad, "model_feature_linear": feat.grad, "model_views_linears_0": prev.grad, "model_rgb_linear": rgb.grad})
    # sum |terms| of every gradient entry: the scale its rounding error is relative to
    bound = {}
    for name, d in dy.items():
        bound[name + ".weight"] = float((d.abs().t() @ xin[name].detach().abs()).max())
        bound[name + ".bias"] = float(d.abs().sum(0).max())
    return dict(out=out.detach(), grads={k: v.grad for k, v in pd.items()}, bound=bound, pts=pts, views=views, hs=hs, pre=pre, feat=feat, prev=prev, hv=hv,
                pd=pd)
