"""Layer-by-layer diagnosis of the classic-NeRF training kernels on a B200: decodes the scratch records of nrf_mlp_nerf_fwd_train /
nrf_mlp_nerf_bwd and compares every activation, every pre-activation gradient and every parameter gradient with an fp64 autograd
evaluation of oracle/restate.py:nerf_forward.  Usage: python scripts/debug/nerf_bwd_check.py [n]"""
import math
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path[:0] = [str(ROOT), str(ROOT / "oracle"), str(ROOT / "tests")]
import restate as O  # noqa: E402
from nerfpp_b200 import ops  # noqa: E402

SAVE_TILE, GRAD_TILE = 684032, 626688


def region(buf, tile_bytes, off, cols, n):
    t = buf.view(-1, tile_bytes)[:, off:off + 256 * cols].contiguous().view(torch.bfloat16)
    t = t.view(-1, 2, cols // 8, 64, 8).permute(0, 1, 3, 2, 4).reshape(-1, cols)
    return t[:n].double()


def rel(a, b):
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-300))


def main(n=1000, gain=1.0):
    torch.manual_seed(0)
    from test_gpu_mlp_nerf import _inputs, _params, _xavier_params, _train_reference as reference
    p = _params(seed=3) if gain == 1.0 else _xavier_params(seed=3, gain=gain)
    x = _inputs(n, seed=4)
    gout = torch.randn(n, 4, dtype=torch.float64, device="cuda") * 1e-3
    R = reference(x, p, gout, False)
    E = reference(x, p, gout, True)
    pd, pts, views, hs, pre, feat, prev, hv, out_ref = (R[k] for k in ("pd", "pts", "views", "hs", "pre", "feat", "prev", "hv", "out"))
    gout = gout.double()

    packed = ops.mlp_nerf_pack(p, train=True)
    out, saved = ops.mlp_nerf_fwd_train(packed, x)
    torch.cuda.synchronize()
    print(f"fwd_train out rel err {rel(out.double(), out_ref.detach()):.3e}")
    print(f"  saved pts   {rel(region(saved, SAVE_TILE, 0, 64, n)[:, :63], pts):.3e}")
    print(f"  saved views {rel(region(saved, SAVE_TILE, 16384, 32, n)[:, :27], views):.3e}")
    for l in range(1, 9):
        print(f"  saved h{l}    {rel(region(saved, SAVE_TILE, 24576 + (l - 1) * 65536, 256, n), hs[l].detach()):.3e}")
    print(f"  saved feat  {rel(region(saved, SAVE_TILE, 548864, 256, n), feat.detach()):.3e}")
    print(f"  saved hv    {rel(region(saved, SAVE_TILE, 614400, 128, n), hv.detach()):.3e}")

    grads = {k: torch.zeros_like(v) for k, v in p.items()}
    ws = torch.zeros((n + 127) // 128 * GRAD_TILE, dtype=torch.uint8, device="cuda")
    ops.mlp_nerf_bwd(packed, saved, gout.float().contiguous(), grads, workspace=ws)
    torch.cuda.synchronize()
    print(f"  dOut  {rel(region(ws, GRAD_TILE, 0, 16, n)[:, :4], gout):.3e}")
    print(f"  d_hv  {rel(region(ws, GRAD_TILE, 4096, 128, n), prev.grad):.3e}")
    ours, ref = region(ws, GRAD_TILE, 4096, 128, n), prev.grad
    unm = gout[:, :3] @ pd["model_rgb_linear.weight"].detach()
    print(f"    d_hv vs unmasked ref {rel(ours, unm):.3e}; zeros ours {float((ours == 0).double().mean()):.3f} ref {float((ref == 0).double().mean()):.3f}; "
          f"ours!=0&ref==0 {int(((ours != 0) & (ref == 0)).sum())}  ours==0&ref!=0 {int(((ours == 0) & (ref != 0)).sum())}")
    both = (ours != 0) & (ref != 0)
    print(f"    on common support: {float(((ours - ref).abs() * both).max() / ref.abs().max()):.3e}")
    bad = ((ours - ref).abs() > 0.02 * ref.abs().max())
    print("    bad per 8-col chunk:", bad.view(n, 16, 8).sum((0, 2)).tolist())
    rb = torch.zeros(4, dtype=torch.long)
    for r4 in range(4):
        idx = (torch.arange(n, device="cuda") % 128) // 32 == r4
        rb[r4] = int(bad[idx].sum())
    print("    bad per 32-row block of the tile:", rb.tolist(), " first bad (row, col):", torch.nonzero(bad)[:8].tolist())
    i, j = (torch.nonzero(bad)[0].tolist() if bad.any() else (0, 0))
    print(f"    sample ({i},{j}): ours {float(ours[i, j]):.4e} ref {float(ref[i, j]):.4e} unmasked {float(unm[i, j]):.4e} hv {float(hv[i, j]):.4e}")
    print(f"  d_feat {rel(region(ws, GRAD_TILE, 36864, 256, n), feat.grad):.3e}")
    for l in range(7, -1, -1):
        print(f"  dY_{l}  {rel(region(ws, GRAD_TILE, 102400 + l * 65536, 256, n), pre[l].grad):.3e}")
    print(f"  vs emulation: d_hv {rel(region(ws, GRAD_TILE, 4096, 128, n), E['prev'].grad):.3e}  d_feat {rel(region(ws, GRAD_TILE, 36864, 256, n), E['feat'].grad):.3e}  "
          + "  ".join(f"dY_{l} {rel(region(ws, GRAD_TILE, 102400 + l * 65536, 256, n), E['pre'][l].grad):.3e}" for l in range(7, -1, -1)))
    for k in p:
        g, r, e = grads[k].double(), pd[k].grad, E["pd"][k].grad
        cos = float((g * r).sum() / (g.norm() * r.norm()).clamp_min(1e-300))
        print(f"  grad {k:32s} vs fp64 {rel(g, r):.3e} (L2 {float((g - r).norm() / r.norm()):.3e}, cos {cos:.5f})   vs emulation {rel(g, e):.3e} (of sum|terms| {float((g - e).abs().max()) / E["bound"][k]:.2e})   |ref|max {float(r.abs().max()):.3e}")


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 1000, float(sys.argv[2]) if len(sys.argv) > 2 else 1.0)
