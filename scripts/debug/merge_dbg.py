import sys
sys.path[:0] = ['.', 'oracle', 'tests']
import torch
from nerfpp_b200 import ops
from nerfpp_b200.pipeline import HashNeRF
BBOX = (-1.5, -1.5, -1.5, 1.5, 1.5, 1.5)
m = HashNeRF(BBOX, log2_hashmap_size=14, seed=3)
with torch.no_grad():
    m.params[:m.n_table] = torch.rand(m.n_table, device="cuda") * 2 - 1
    off = m.n_table
    for fo, fi in ((64, 32), (16, 64), (64, 31), (64, 64), (3, 64)):
        m.params[off:off + fo * fi] = torch.randn(fo * fi, device="cuda") * (2.0 / fi) ** 0.5
        off += fo * fi
m.refresh()
g = torch.Generator().manual_seed(0)
n = 24
o = torch.tensor([0.3, -0.2, 4.0]).repeat(n, 1) + 0.05 * torch.randn(n, 3, generator=g)
d = torch.tensor([0.0, 0.0, -1.0]) + 0.25 * torch.randn(n, 3, generator=g)
o, d = o.cuda(), d.cuda()
rb = ops.rays_prepare(o, d, BBOX, 0.0, True)
sh = ops.sh_encode(rb[:, 8:11], 4)
z = ops.z_sample(rb, m.t_vals)
_, _, _, raw = m._network(rb, z, sh)
coarse = ops.composite_fwd(raw, z, d, False)
merged, zs = ops.sample_pdf_merge(z, coarse["weights"], m.u, want_samples=True)
print("nan merged", int(torch.isnan(merged).sum()), "nan zs", int(torch.isnan(zs).sum()), "nan w", int(torch.isnan(coarse["weights"]).sum()))
bad = ((merged[:, 1:] < merged[:, :-1]).any(-1)).nonzero().flatten().tolist()
print("unsorted rows", bad)
for r in bad[:3]:
    print("row", r, "near/far", rb[r, 6:8].tolist())
    print(" z sorted", bool((z[r, 1:] >= z[r, :-1]).all()), " zs sorted", bool((zs[r, 1:] >= zs[r, :-1]).all()))
    w = coarse["weights"][r]
    print(" w min/max/sum", float(w.min()), float(w.max()), float(w.sum()))
    i = (merged[r, 1:] < merged[r, :-1]).nonzero().flatten().tolist()
    print(" inversion idx", i[:10], [merged[r, k - 1:k + 3].tolist() for k in i[:3]])
    print(" zs", zs[r].tolist()[:8], "...", zs[r].tolist()[-8:])
    ref = torch.sort(torch.cat([z[r], zs[r]])).values
    print(" vs sort: max diff", float((ref - merged[r]).abs().max()))
