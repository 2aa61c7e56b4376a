"""Which rows of the fine pass differ between the three reuse modes (debug aid for tests/test_gpu_pipeline.py)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from nerfpp_b200.pipeline import HashNeRF, synthetic_rays  # noqa: E402

BBOX = [-1.5, -1.5, -1.5, 1.5, 1.5, 1.5]
m = HashNeRF(BBOX, log2_hashmap_size=14, seed=3)
g = torch.Generator(device="cuda").manual_seed(1234)
with torch.no_grad():
    m.params[:m.n_table] = torch.rand(m.n_table, device="cuda", generator=g) * 2 - 1
    m.params[m.n_table:] = torch.randn(m.params.numel() - m.n_table, device="cuda", generator=g) * 0.2
m.refresh()
o, d, _ = synthetic_rays(4096, seed=4)
outs = []
for rows, raw in ((False, False), (True, False), (True, True)):
    m.reuse_coarse_rows, m.reuse_coarse_raw = rows, raw
    outs.append(m.render_rays(o, d, keep_for_backward=True))
a = outs[0]
for n, b in zip(("rows", "rows+raw"), outs[1:]):
    for i, name in ((1, "enc"), (2, "keep"), (3, "raw")):
        x, y = a["_saved"][i], b["_saved"][i]
        x, y = x.reshape(4096 * 192, -1), y.reshape(4096 * 192, -1)
        bad = (x != y).any(dim=1).nonzero().flatten()
        print(n, name, "differing rows:", bad.numel(), bad[:10].tolist())
        if bad.numel():
            r = int(bad[0])
            print("  row", r, "ray", r // 192, "pos", r % 192, x[r].tolist()[:4], y[r].tolist()[:4])
    print(n, "rgb equal:", torch.equal(a["rgb"], b["rgb"]), "z equal:", torch.equal(a["z"], b["z"]))
