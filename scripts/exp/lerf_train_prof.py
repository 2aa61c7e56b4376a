"""One LeRFField training configuration (BASELINE C5 shape, 1024 rays) stepped eagerly a few times: the command profiled with ncu
(launch list / --set full of the LeRF training kernels).  Prints per-entry CUDA-event times when run without a profiler."""
import sys, json, torch
sys.path.insert(0, ".")
from nerfpp_b200 import ops
from nerfpp_b200.lerf import LeRFField
from nerfpp_b200.pipeline import synthetic_rays

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
f = LeRFField(seed=0, lr=5e-4)
g = torch.Generator().manual_seed(0)
for v in f.weights.values():
    v.copy_((torch.randn(v.shape, generator=g) * (2.0 / v.shape[1]) ** 0.5).cuda())
f.params[:f.n_table].copy_((torch.rand(f.n_table, generator=g) * 2 - 1).cuda())
f.refresh()
o, d, _ = synthetic_rays(1024, seed=5)
tgt = torch.nn.functional.normalize(torch.randn(1024, 512, generator=g), dim=-1).cuda()
for _ in range(2):
    f.train_step(o, d, tgt)
torch.cuda.synchronize()
t = ops.KernelTimer()
ops.set_timer(t)
for _ in range(steps):
    f.train_step(o, d, tgt)
torch.cuda.synchronize()
ops.set_timer(None)
print(json.dumps({"loss": float(f.loss), "ms_per_entry": {k: round(ms / n, 4) for k, (n, ms) in sorted(t.summary().items(), key=lambda kv: -kv[1][1])}}))
