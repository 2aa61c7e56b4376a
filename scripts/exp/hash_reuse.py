"""Experiment: cost of the fine-pass hash encode with / without coarse-row reuse, and of fused point generation."""
import sys, torch
sys.path.insert(0, ".")
from nerfpp_b200 import ops
from nerfpp_b200.pipeline import HashNeRF, synthetic_rays

m = HashNeRF()
o, d, t = synthetic_rays(4096, seed=1)
rb = ops.rays_prepare(o, d, m.bbox, 0.0, True)
sh = ops.sh_encode(rb[:, 8:11], 4)
z = ops.z_sample(rb, m.t_vals)
enc_c, keep_c, raw = m._network(rb, z, sh)
co = ops.composite_fwd(raw, z, d)
zf, src = ops.sample_pdf_merge(z, co["weights"], m.u, want_perm=True)
print("claimed fraction", float((src[:, 128:] >= 0).float().mean()))
pts = ops.sample_points(rb, zf)

def timeit(f, n=20):
    for _ in range(3): f()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): f()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e3

none = None
print("fine points-array      us", timeit(lambda: ops.hash_encode_fwd(m.grid, m.table_f16, pts.view(-1, 3), out_f16=True)))
print("fine rays no reuse     us", timeit(lambda: ops.hash_encode_rays_fwd(m.grid, m.table_f16, rb, zf)))
print("fine rays reuse        us", timeit(lambda: ops.hash_encode_rays_fwd(m.grid, m.table_f16, rb, zf, reuse=(src, enc_c, keep_c, 64))))
pc = ops.sample_points(rb, z)
print("coarse points-array    us", timeit(lambda: ops.hash_encode_fwd(m.grid, m.table_f16, pc.view(-1, 3), out_f16=True)))
print("coarse rays            us", timeit(lambda: ops.hash_encode_rays_fwd(m.grid, m.table_f16, rb, z)))
zz = zf[:, :128].contiguous()
print("128/ray rays           us", timeit(lambda: ops.hash_encode_rays_fwd(m.grid, m.table_f16, rb, zz)))
