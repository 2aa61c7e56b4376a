"""Generates the CUDA golden fixtures by executing the reference's own CUDA kernels (CuHashEmbedder fwd/bwd,
CuSHEncoder; oracle/_ref/nerfpp_ref_cuda.so, built from /root/reference/src with nvcc -arch=sm_100) on a B200:

    gpurun -- 'python tests/golden/make_golden_cuda.py gpurun_out/golden'

then the .npz files are copied from gpurun_out/golden/ into tests/golden/ and committed.  They pin the CUDA-only
reference stages for machines where the reference module cannot be loaded.
"""
import sys
from pathlib import Path

import numpy as np
import torch

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent.parent / "oracle" / "_ref"))
sys.path.insert(0, str(HERE.parent.parent))
import nerfpp_ref_cuda as R  # noqa: E402

out_dir = Path(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/golden")
out_dir.mkdir(parents=True, exist_ok=True)
BBOX = torch.tensor([-1.5, -1.5, -1.5, 1.5, 1.5, 1.5]).cuda()   # BoundingBox is a plain member: ->to(device) does not move it


def npz(name, **kw):
    np.savez_compressed(out_dir / name, **{k: (v.detach().cpu().numpy() if isinstance(v, torch.Tensor) else np.asarray(v)) for k, v in kw.items()})
    print("wrote", out_dir / name)


def cuhash():
    R.manual_seed(42)
    torch.manual_seed(42)
    pipe = R.make_cuhash(BBOX, 16, 2, 10, 16, 512, 4, 2, 64, 15, 3, 64)     # small table (2^10 per level) for fixture size
    bufs = dict(zip(pipe.embed_buffer_names(), pipe.embed_buffers()))
    table = pipe.embed_params()[0]
    with torch.no_grad():
        table.copy_((torch.rand_like(table) * 2 - 1).half().float())
    pts = torch.rand(768, 3, device="cuda") * 3 - 1.5
    pts[0] = -1.5
    pts[1] = 1.5
    pts[2] = 0.0
    pts[3] = torch.tensor([1.7, 0.1, 0.2])
    pts[4] = torch.tensor([0.1, -9.0, 0.2])
    enc, keep = pipe.embed(pts)
    g = (torch.randn_like(enc) * 1e-2)
    g[5] = 0
    enc.backward(g)
    npz("cuhash.npz", table_f16=table.detach().half(), primes=bufs["embedder_primes"], biases=bufs["embedder_biases"],
        feat_local_idx=bufs["embedder_feat_local_idx"], feat_local_size=bufs["embedder_feat_local_size"],
        points=pts, enc=enc, keep=keep, grad_enc=g, grad_table=table.grad)


def cush():
    torch.manual_seed(1)
    d = torch.randn(96, 3, device="cuda")
    d[:48] = torch.nn.functional.normalize(d[:48], dim=-1)
    kw = {f"sh{deg}": R.cu_sh_encoder(d.contiguous(), deg) for deg in range(1, 9)}
    npz("cush.npz", dirs=d, **kw)


def level_scales():
    from nerfpp_b200 import ops
    npz("level_scales.npz", s_16_512_16=ops.hash_level_scales(16, 512, 16, "cuda"), s_16_1024_16=ops.hash_level_scales(16, 1024, 16, "cuda"),
        s_8_128_8=ops.hash_level_scales(8, 128, 8, "cuda"))


def full_render():
    """RenderRays of the reference's CUDA instantiation NeRFRenderer<CuHashEmbedder,CuSHEncoder,NeRFSmall> (single chunk)."""
    R.manual_seed(7)
    torch.manual_seed(7)
    pipe = R.make_cuhash(BBOX, 16, 2, 12, 16, 512, 4, 2, 64, 15, 3, 64)
    bufs = dict(zip(pipe.embed_buffer_names(), pipe.embed_buffers()))
    table = pipe.embed_params()[0]
    with torch.no_grad():
        table.copy_((torch.rand_like(table) * 2 - 1).half().float())
        for p in pipe.model_params():
            p.copy_(torch.randn_like(p) * (2.0 / p.shape[1]) ** 0.5)
    n = 48
    o = torch.tensor([0.3, -0.2, 4.0], device="cuda").repeat(n, 1) + 0.05 * torch.randn(n, 3, device="cuda")
    d = torch.tensor([0.0, 0.0, -1.0], device="cuda") + 0.25 * torch.randn(n, 3, device="cuda")
    res = pipe.render(o, d, 64, 128, 4096, False, True)
    kw = {f"w{i}": p for i, p in enumerate(pipe.model_params())}
    npz("cuhash_render.npz", table_f16=table.detach().half(), primes=bufs["embedder_primes"], o=o, d=d, rgb=res["rgb"], depth=res["depth"],
        acc=res["acc"], disp=res["disp"], weights=res["weights"], **kw)


if __name__ == "__main__":
    cuhash()
    cush()
    level_scales()
    full_render()
