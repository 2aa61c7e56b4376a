"""Quantisation-point-exact emulation of the fused NeRFSmall kernels (test infrastructure).

The reference MLP (src/NeRF.cpp:363-412) is fp32 SGEMM.  The sm_100a kernels run the forward chain (and its recompute in
the backward) in fp16 x fp16 and the gradient chain in bf16 x bf16, all with fp32 accumulation, re-quantising each
activation / gradient where it becomes a tensor-core operand (for the dW products the activations are packed to bf16 straight
from the fp32 accumulators; encodings and view channels, which arrive as fp16 values, are re-quantised).
Against the fp32 reference such a chain is only comparable in a norm sense: a pre-activation within rounding error of zero
takes the other ReLU branch and changes that row's gradient by a whole term, which a max-norm bound cannot absorb.  So
parity is established in four steps (tests/test_gpu_mlp.py):
    kernel == this emulation                      tight, max-norm (catches every indexing / layout / masking bug)
    kernel ~  fp32 reference, forward             max-norm rel <= 1e-2 (the bf16 tolerance class of the north star)
    kernel ~  fp32 arithmetic on the SAME active sets (fp32_with_masks), gradients: max-norm rel <= 1.5e-2
    active sets vs the fp32 reference's           differing ReLU units < 2 %, median per-row gradient error <= 1e-2
"""
import torch


def bf(x):
    return x.to(torch.bfloat16).to(torch.float32)


def hf(x):
    return x.to(torch.float16).to(torch.float32)


def pad_w2(w2):
    """[64,31] -> [64,32]: columns [views 16 | sigma slot (zero) | geo 15], the kernel's A2 layout."""
    return torch.cat([w2[:, :16], torch.zeros(w2.shape[0], 1), w2[:, 16:]], -1)


def forward_backward(ws, x, g, keep=None):
    """ws: 5 fp32 weights [out,in]; x [N,48] = [enc(32, fp16 values) | views(16)]; g [N,4] upstream gradient.
    Returns out [N,4], gx [N,48], [dW0..dW4] evaluated with the kernels' rounding points."""
    w0, w1, w2, w3, w4 = [w.float() for w in ws]
    w2p = pad_w2(w2)
    x0 = hf(x[:, :32])
    views = hf(x[:, 32:])
    acc0 = x0 @ hf(w0).t()
    a1 = hf(torch.relu(acc0))
    x1b = bf(torch.relu(acc0))       # the dW copy of an activation is packed to bf16 straight from the fp32 accumulator
    d1 = a1 @ hf(w1).t()
    sigma = d1[:, 0].clone()
    d1z = d1.clone()
    d1z[:, 0] = 0.0
    a2 = torch.cat([views, hf(d1z)], -1)
    acc2 = a2 @ hf(w2p).t()
    a3 = hf(torch.relu(acc2))
    x3b = bf(torch.relu(acc2))
    acc3 = a3 @ hf(w3).t()
    a4 = hf(torch.relu(acc3))
    x4b = bf(torch.relu(acc3))
    c = a4 @ hf(w4).t()
    if keep is not None:
        sigma = sigma * keep.float()
    out = torch.cat([c, sigma[:, None]], -1)

    g = g.float().clone()
    if keep is not None:
        g[:, 3] = g[:, 3] * keep.float()
    d4 = bf(g[:, :3])
    dA4 = d4 @ bf(w4)
    dD3 = bf(dA4 * (a4 > 0))
    dA3 = dD3 @ bf(w3)
    dD2 = bf(dA3 * (a3 > 0))
    dA2 = dD2 @ bf(w2p)
    dviews = dA2[:, :16]
    dd1_f = dA2[:, 16:].clone()
    dd1_f[:, 0] += g[:, 3]
    dd1 = bf(dd1_f)
    dA1 = dd1 @ bf(w1)
    dD0 = bf(dA1 * (a1 > 0))
    denc = dD0 @ bf(w0)
    gx = torch.cat([denc, dviews], -1)
    dW0 = dD0.t() @ bf(x0)
    dW1 = dd1.t() @ x1b
    dW2p = dD2.t() @ bf(a2)
    dW2 = torch.cat([dW2p[:, :16], dW2p[:, 17:]], -1)
    dW3 = dD3.t() @ x3b
    dW4 = d4.t() @ x4b
    return out, gx, [dW0, dW1, dW2, dW3, dW4], (a1 > 0, a3 > 0, a4 > 0)


def fp32_with_masks(ws, x, g, masks, keep=None):
    """NeRFSmall forward/backward in plain fp32 (src/NeRF.cpp:363-412 and its autograd adjoint) with the ReLU active
    sets given instead of derived: relu(h) := h * mask.  With the fp32 reference's own masks this IS the reference."""
    w0, w1, w2, w3, w4 = [w.float() for w in ws]
    m1, m3, m4 = [m.float() for m in masks]
    x0, views = x[:, :32].float(), x[:, 32:].float()
    a1 = (x0 @ w0.t()) * m1
    d1 = a1 @ w1.t()
    a2 = torch.cat([views, d1[:, 1:]], -1)
    a3 = (a2 @ w2.t()) * m3
    a4 = (a3 @ w3.t()) * m4
    g = g.float().clone()
    if keep is not None:
        g[:, 3] = g[:, 3] * keep.float()
    d4 = g[:, :3]
    dD3 = (d4 @ w4) * m4
    dD2 = (dD3 @ w3) * m3
    dA2 = dD2 @ w2
    dd1 = torch.cat([g[:, 3:4], dA2[:, 16:]], -1)
    dD0 = (dd1 @ w1) * m1
    gx = torch.cat([dD0 @ w0, dA2[:, :16]], -1)
    return gx, [dD0.t() @ x0, dd1.t() @ a1, dD2.t() @ a2, dD3.t() @ a3, d4.t() @ a4]


def fp32_masks(ws, x):
    w0, w1, w2, w3, w4 = [w.float() for w in ws]
    h0 = x[:, :32].float() @ w0.t()
    d1 = torch.relu(h0) @ w1.t()
    h2 = torch.cat([x[:, 32:].float(), d1[:, 1:]], -1) @ w2.t()
    h3 = torch.relu(h2) @ w3.t()
    return h0 > 0, h2 > 0, h3 > 0


def rel_l2(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))
