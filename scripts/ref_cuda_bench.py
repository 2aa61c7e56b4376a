"""Same-box GPU comparison: the reference's own CUDA instantiation
NeRFRenderer<CuHashEmbedder,CuSHEncoder,NeRFSmall> + huber + backward + torch::optim::Adam (oracle/_ref/nerfpp_ref_cuda.so,
built unmodified from /root/reference/src with nvcc -arch=sm_100) timed on one B200 at the C2 shape.
TEST/BENCH INFRASTRUCTURE: prints one JSON line; never imported by the product path.

    python scripts/ref_cuda_bench.py [rays] [steps] [ref|dropin]

`dropin` runs the SAME C++ training loop (bindings: Render + huber + backward + torch::optim::Adam, NeRFExecutor::Train's
lines) on this repo's C++ drop-in classes (nerfpp_b200/lib/nerfpp_b200_torch.so) instead of the reference's.
"""
import json
import os
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "oracle" / "_ref")]
import torch  # noqa: E402
from nerfpp_b200.pipeline import synthetic_rays  # noqa: E402

rays = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 30
which = sys.argv[3] if len(sys.argv) > 3 else "ref"
if which == "ref":
    import nerfpp_ref_cuda as R  # noqa: E402
else:
    from nerfpp_b200 import build  # noqa: E402
    sys.path.insert(0, str(build.build_host().parent))
    import nerfpp_b200_torch as R  # noqa: E402
R.manual_seed(42)
saved = os.dup(1)
devnull = os.open(os.devnull, os.O_WRONLY)
os.dup2(devnull, 1)
try:
    pipe = R.make_cuhash(torch.tensor([-1.5, -1.5, -1.5, 1.5, 1.5, 1.5]).cuda(), 16, 2, 19, 16, 512, 4, 2, 64, 15, 3, 64)
    pipe.init_model()
finally:
    os.dup2(saved, 1)
if which == "dropin-fused":
    pipe.use_fused_adam(True)
o, d, tgt = synthetic_rays(rays, device="cuda", seed=0)
pipe.train_steps(o, d, tgt, 5, 64, 128, rays, True, 1e-2, 250)   # warm-up; single chunk
torch.cuda.synchronize()
secs, losses = pipe.train_steps(o, d, tgt, steps, 64, 128, rays, True, 1e-2, 250)
secs = sorted(secs)
med = secs[len(secs) // 2]
print(json.dumps({"impl": {"ref": "reference-cuda", "dropin": "nerfpp_b200 C++ drop-in classes in the reference's loop (torch::optim::Adam)",
                           "dropin-fused": "nerfpp_b200 C++ drop-in classes + FusedAdam in the reference's loop"}[which], "metric": "train_rays_per_s", "value": rays / med, "unit": "rays/s", "rays": rays,
                  "steps": steps, "ms_per_step_median": med * 1e3, "ms_per_step_mean": sum(secs) / len(secs) * 1e3,
                  "timing": "steady_clock around each step incl. loss.item() sync (reference's own method)", "loss_last": losses[-1]}))
