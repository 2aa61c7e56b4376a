import sys, torch
from pathlib import Path
ROOT = Path(".").resolve()
sys.path[:0] = [str(ROOT), str(ROOT / "oracle" / "_ref"), str(ROOT / "nerfpp_b200" / "lib")]
import nerfpp_b200_torch as host
import nerfpp_ref_cuda as ref
BBOX = [-1.5, -1.5, -1.5, 1.5, 1.5, 1.5]
pipes = []
for mod, extra in ((host, ()), (ref, (True,))):
    mod.manual_seed(9); torch.manual_seed(9)
    p = mod.make_classic(torch.tensor(BBOX).cuda(), 10, 4, 8, 256, True, *extra); p.init_model(); pipes.append(p)
ours, rf = pipes
K = torch.tensor([[30.0, 0, 12.0], [0, 30.0, 10.0], [0, 0, 1]]).cuda(); c2w = torch.eye(4); c2w[2, 3] = 4.0; c2w = c2w.cuda()
a = ours.render_image(20, 24, K, c2w, 64, 128, 4096, False, True)
b = rf.render_image(20, 24, K, c2w, 64, 128, 4096, False, True)
for k in ("rgb", "acc", "depth", "disp"):
    d = (a[k] - b[k]).abs()
    print(k, a[k].shape, "nan ours", int(torch.isnan(a[k]).sum()), "nan ref", int(torch.isnan(b[k]).sum()), "max abs diff", float(d[~torch.isnan(d)].max()), "ref max", float(b[k][~torch.isnan(b[k])].abs().max()))
i = torch.argmax(torch.nan_to_num((a["rgb"] - b["rgb"]).abs().reshape(-1, 3).sum(-1)))
print("worst pixel", int(i), a["rgb"].reshape(-1, 3)[i], b["rgb"].reshape(-1, 3)[i], a["acc"].reshape(-1)[i], b["acc"].reshape(-1)[i])
