"""Loss curves of the classic NeRF training loop (reference initialisation, same seeds) on the fused bf16 tcgen05 kernels vs torch::linear fp32:
python scripts/debug/classic_curve.py [steps] -> one JSON line with the losses every 25 steps."""
import json
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path[:0] = [str(ROOT), str(ROOT / "nerfpp_b200" / "lib")]
import nerfpp_b200_torch as host  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 500
bbox = torch.tensor([-1.5, -1.5, -1.5, 1.5, 1.5, 1.5]).cuda()
g = torch.Generator().manual_seed(0)
rays = 1024
d = torch.nn.functional.normalize(torch.randn(rays, 3, generator=g), dim=-1)
o = (-4.0 * d + 0.3 * torch.randn(rays, 3, generator=g)).cuda()
d = d.cuda()
# a smooth target image-like function of the ray (not noise), so that there is something to learn
tgt = (0.5 + 0.5 * torch.sin(3.0 * torch.stack([d[:, 0] + o[:, 1], d[:, 1] - o[:, 2], d[:, 2] * 2], -1))).contiguous()
res = {}
for name, fused in (("fused_bf16", True), ("aten_fp32", False)):
    host.manual_seed(1)
    torch.manual_seed(1)
    p = host.make_classic(bbox, 10, 4, 8, 256, True)
    p.init_model()
    host.classic_set_fused_training(p, fused)
    host.classic_set_fused_embedding(p, fused)
    _, losses = p.train_steps(o, d, tgt, steps, 64, 128, 1 << 20, True, 5e-4, 250)
    res[name] = [round(l, 6) for l in losses[::25]] + [round(losses[-1], 6)]
print(json.dumps({"steps": steps, "every": 25, **res}))
