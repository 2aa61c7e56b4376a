"""Quantisation-point emulation of the fused LeRF head (nerfpp_b200/csrc/lerf_tc.cu) — test infrastructure.

The reference head (src/LeRF.cpp:28-111) is four fp32 SGEMMs.  The sm_100a kernel runs them as fp16 x fp16 tensor-core products with fp32
accumulation and re-quantises every activation where it becomes the next layer's A operand (h1, geo, h2); the norm of the un-formed embedding
comes from G = W_e1^T W_e1, summed in fp32, divided by a power of two so that its largest diagonal entry is <= 256, and rounded to fp16.
This file restates those rounding points in torch so that their cost can be bounded on CPU against the fp64 oracle
(tests/test_lerf_precision.py); summation ORDER inside a dot product is not emulated (fp32 accumulation, 1e-7 class).
"""
import torch


def hf(x):
    return x.to(torch.float16).to(torch.float32)


def g_scale(w_e1: torch.Tensor) -> float:
    """lerf_gscale_kernel: the smallest power of two s >= 1 with max_k (W^T W)_kk / s <= 256."""
    mx = float((w_e1.float() ** 2).sum(0).max())
    s = 1.0
    while mx > 256.0 * s and s < 1e30:
        s *= 2.0
    return s


def forward(x: torch.Tensor, sigma_w, le_w):
    """x [N,128] (fp16-representable); weights fp32 [out,in].  Returns the HIDDEN program's outputs: (sigma [N], h2 [N,256] fp16 values, q [N])
    and the RAW program's raw_le [N,513]."""
    w_s0, w_s1 = (hf(w.float()) for w in sigma_w)
    w_e0, w_e1 = (hf(w.float()) for w in le_w)
    x = hf(x.float())
    h1 = hf(torch.relu(x @ w_s0.t()))
    s = h1 @ w_s1.t()
    sigma, geo = s[:, 0], hf(s[:, 1:])
    h2 = hf(torch.relu(torch.cat([geo, x], -1) @ w_e0.t()))
    scale = g_scale(le_w[1])
    g16 = hf((le_w[1].float().t() @ le_w[1].float()) / scale)
    q = ((h2 @ g16.t()) * h2).sum(-1) * scale
    e = h2 @ w_e1.t()
    raw = torch.cat([e / e.norm(dim=-1, keepdim=True).clamp_min(1e-8), sigma[:, None]], -1)
    return sigma, h2, q, raw
