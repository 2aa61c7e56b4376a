"""A few eager training steps at the BASELINE C2 shape: the command profiled under ncu for single-kernel captures."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from nerfpp_b200.pipeline import HashNeRF, synthetic_rays  # noqa: E402

m = HashNeRF((-1.5, -1.5, -1.5, 1.5, 1.5, 1.5), seed=42)
o, d, t = synthetic_rays(4096, seed=1)
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 4):
    m.train_step(o, d, t)
torch.cuda.synchronize()
print("loss", float(m.loss))
