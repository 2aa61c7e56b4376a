"""Render leg (1920x1080, 64+64 samples) vs ray-chunk size."""
import math
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[2]))
from nerfpp_b200.pipeline import HashNeRF  # noqa: E402

m = HashNeRF((-1.5, -1.5, -1.5, 1.5, 1.5, 1.5), seed=0)
H, W = 1080, 1920
focal = 0.5 * W / math.tan(0.5 * 0.6911)
K = [[focal, 0, 0.5 * W], [0, focal, 0.5 * H], [0, 0, 1]]
c2w = torch.eye(4)
c2w[2, 3] = 4.0
for chunk in (32768, 65536, 131072, 262144, 524288):
    m.render_image(H, W, K, c2w, chunk=chunk, n_importance=64)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        m.render_image(H, W, K, c2w, chunk=chunk, n_importance=64)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    print(f"chunk {chunk:7d}: {ms:7.2f} ms/frame  {H * W * 192 / ms / 1e3:8.1f} Msamples/s  peak mem {torch.cuda.max_memory_allocated() / 2**30:.2f} GiB", flush=True)
