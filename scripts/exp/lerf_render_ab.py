"""C5 field, C4 frame: RenderedLangEmbedding of 135 image rows (1/8 frame) under NRF_RENDER_RAY_GROUP; prints ms and per-kernel times."""
import json
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from nerfpp_b200 import ops  # noqa: E402
from nerfpp_b200.lerf import LeRFField  # noqa: E402

BBOX = [-1.5, -1.5, -1.5, 1.5, 1.5, 1.5]
field = LeRFField(BBOX, seed=0)
g = torch.Generator().manual_seed(0)
for v in field.weights.values():
    v.copy_((torch.randn(v.shape, generator=g) * (2.0 / v.shape[1]) ** 0.5).cuda())
field.table.copy_((torch.rand(field.n_table, generator=g) * 2 - 1).cuda())
field.refresh()
H, W = 1080, 1920
focal = 0.5 * W / math.tan(0.5 * 0.6911)
K = [[focal, 0, 0.5 * W], [0, focal, 0.5 * H], [0, 0, 1]]
c2w = torch.eye(4)
c2w[2, 3] = 4.0
r0, r1 = 472, 607
import time
use_graph = os.environ.get("LERF_GRAPH", "1") != "0"
field.render_image(H, W, K, c2w, chunk=8192, row_begin=r0, row_end=r1, use_graph=use_graph)      # warm-up at the timed size (allocator, capture)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter()
e0.record()
for _ in range(2):
    field.render_image(H, W, K, c2w, chunk=8192, row_begin=r0, row_end=r1, use_graph=use_graph)
e1.record()
t_host = (time.perf_counter() - t0) * 1e3 / 2
torch.cuda.synchronize()
kt = ops.KernelTimer()
ops.set_timer(kt)
field.render_image(H, W, K, c2w, chunk=8192, row_begin=r0, row_end=r0 + 17, use_graph=False)
ops.set_timer(None)
torch.cuda.synchronize()
print(json.dumps({"env": {k: v for k, v in os.environ.items() if k.startswith("NRF_")}, "graph": use_graph, "ms_eighth_frame": e0.elapsed_time(e1) / 2, "ms_host_enqueue": t_host,
                  "kernels_ms_per_launch": {k: round(ms / n, 4) for k, (n, ms) in kt.summary().items()}}))
