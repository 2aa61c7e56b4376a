"""BASELINE C1 network in the reference's own training loop (NeRFExecutor::Train's lines: Render + huber + backward + Adam) on the C++
drop-in classes NeRFRenderer<Embedder,Embedder,NeRF>: fused tcgen05 training kernels against the same module on torch::linear +
LibTorch autograd (fp32 cuBLAS, the reference's arithmetic on this GPU).  Prints one JSON line.
Usage: python scripts/classic_train_bench.py [rays] [steps]"""
import json
import statistics
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path[:0] = [str(ROOT), str(ROOT / "nerfpp_b200" / "lib")]
import nerfpp_b200_torch as host  # noqa: E402


def main(rays=1024, steps=20, only=None):
    bbox = torch.tensor([-1.5, -1.5, -1.5, 1.5, 1.5, 1.5]).cuda()
    g = torch.Generator().manual_seed(0)
    d = torch.nn.functional.normalize(torch.randn(rays, 3, generator=g), dim=-1)
    o = (-4.0 * d + 0.3 * torch.randn(rays, 3, generator=g)).cuda()
    d = d.cuda()
    tgt = torch.rand(rays, 3, generator=g).cuda()
    res = {}
    for name, fused in (("fused_tcgen05", True), ("aten_fp32", False)):
        if only and name != only:
            continue
        host.manual_seed(1)
        torch.manual_seed(1)
        p = host.make_classic(bbox, 10, 4, 8, 256, True)
        p.init_model()
        host.classic_set_fused_training(p, fused)
        p.use_fused_adam(fused)          # FusedAdam (one nrf_adam_step per parameter tensor) next to the fused MLP; torch::optim::Adam otherwise
        p.train_steps(o, d, tgt, 3, 64, 128, 1 << 20, True, 5e-4, 250)
        torch.cuda.synchronize()
        secs, losses = p.train_steps(o, d, tgt, steps, 64, 128, 1 << 20, True, 5e-4, 250)
        res[name] = {"ms_per_step": 1e3 * statistics.median(secs), "rays_per_s": rays / statistics.median(secs), "loss_first": losses[0], "loss_last": losses[-1]}
    if not only:
        res["speedup"] = res["aten_fp32"]["ms_per_step"] / res["fused_tcgen05"]["ms_per_step"]
    res["config"] = f"C1 network, {rays} rays x (64+128) samples per step, Adam, {steps} steps (median wall clock per step incl. loss.item())"
    print(json.dumps(res))


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 1024, int(sys.argv[2]) if len(sys.argv) > 2 else 20, sys.argv[3] if len(sys.argv) > 3 else None)
