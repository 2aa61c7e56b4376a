"""CPU bound on what the fused LeRF head's rounding points cost against the fp64 oracle (tests/lerf_emul.py restates them): the tolerance the
GPU parity tests use (rel 1e-2, tests/test_gpu_lerf.py) has an order of magnitude of headroom, the fp16 copy of G = W_e1^T W_e1 keeps the
embedding norm to ~1e-3, and the power-of-two scale keeps it finite for a large last layer."""
import math

import pytest
import torch

import lerf_emul as E
import restate as O

SHAPES = ((256, 128), (33, 256), (256, 160), (512, 256))


def _case(seed, last_gain=1.0):
    g = torch.Generator().manual_seed(seed)
    w = [torch.randn(o, i, generator=g) * math.sqrt(2.0 / i) for o, i in SHAPES]
    w[3] = w[3] * last_gain
    x = torch.randn(2048, 128, generator=g).half().float()
    return x, w


@pytest.mark.parametrize("seed,last_gain", [(0, 1.0), (1, 1.0), (2, 200.0), (3, 0.01)])
def test_rounding_points_stay_an_order_of_magnitude_inside_the_gpu_tolerance(seed, last_gain):
    x, w = _case(seed, last_gain)
    sigma, h2, q, raw = E.forward(x, w[:2], w[2:])
    wd = [t.double() for t in w]
    ref = O.lerf_forward(x.double(), wd[:2], wd[2:])
    h1 = torch.relu(x.double() @ wd[0].t())
    sg = h1 @ wd[1].t()
    h2_ref = torch.relu(torch.cat([sg[:, 1:], x.double()], -1) @ wd[2].t())
    q_ref = ((h2_ref @ wd[3].t()) ** 2).sum(-1)
    assert bool(torch.isfinite(q).all()) and bool(torch.isfinite(raw).all())
    assert float((raw[:, :512].double() - ref[:, :512]).abs().max()) <= 2e-3 * float(ref[:, :512].abs().max())
    assert float((sigma.double() - ref[:, 512]).abs().max()) <= 2e-3 * float(ref[:, 512].abs().max())
    assert float((h2.double() - h2_ref).abs().max()) <= 2e-3 * float(h2_ref.abs().max())
    rel_q = ((q.double() - q_ref).abs() / q_ref.clamp_min(1e-30))
    assert float(rel_q.max()) <= 5e-3 and float(rel_q.median()) <= 1e-3
    # the rendered embedding of a 64-sample "ray" through the fused identity vs RenderCLIPEmbedding on the fp64 embedding
    wts = torch.rand(32, 64, generator=torch.Generator().manual_seed(seed)).double()
    c = wts / q.double().sqrt().clamp_min(1e-8).reshape(32, 64)
    fused = torch.nn.functional.normalize((c[..., None] * h2.double().reshape(32, 64, 256)).sum(1) @ wd[3].t(), dim=-1, eps=1e-8)
    exact = O.render_clip_embedding(ref[:, :512].reshape(32, 64, 512), wts[..., None])
    assert float((fused * exact).sum(-1).min()) > 1 - 1e-5


def test_scale_of_the_norm_matrix():
    _, w = _case(5, 200.0)
    s = E.g_scale(w[3])
    diag = float((w[3] ** 2).sum(0).max())
    assert diag > 65504 and s >= 2 and math.log2(s).is_integer() and diag / s <= 256 < 2 * diag / s      # unscaled it overflows fp16
    assert E.g_scale(_case(6)[1][3]) == 1.0                                                              # He-scaled weights: no scaling
