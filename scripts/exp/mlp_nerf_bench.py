"""Experiment: fused classic-NeRF forward (tcgen05) vs the same chain through ATen (cuBLAS fp32, what the reference runs) on one B200."""
import sys, torch
sys.path[:0] = [".", "tests", "oracle"]
import restate as O
from nerfpp_b200 import ops
from test_gpu_mlp_nerf import _params, _inputs

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100 * 100 * 192
p = _params(seed=1)
x = _inputs(n, seed=2)
packed = ops.mlp_nerf_pack(p)

def timeit(f, reps=10):
    for _ in range(3): f()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): f()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps

flop = n * 2.0 * (63 * 256 + 4 * 256 * 256 + 319 * 256 + 2 * 256 * 256 + 256 * 256 + 256 + 283 * 128 + 128 * 3)
t = timeit(lambda: ops.mlp_nerf_fwd(packed, x))
print(f"fused tcgen05 : {t:8.3f} ms  {flop / t / 1e9:8.1f} TFLOP/s useful  ({n} rows)")
with torch.no_grad():
    t2 = timeit(lambda: O.nerf_forward(x, p), reps=3)
    print(f"ATen fp32     : {t2:8.3f} ms  {flop / t2 / 1e9:8.1f} TFLOP/s")
    torch.backends.cuda.matmul.allow_tf32 = True
    t3 = timeit(lambda: O.nerf_forward(x, p), reps=3)
    print(f"ATen tf32     : {t3:8.3f} ms  {flop / t3 / 1e9:8.1f} TFLOP/s")
    ph = {k: v.half() for k, v in p.items()}
    xh = x.half()
    t4 = timeit(lambda: O.nerf_forward(xh, ph), reps=3)
    print(f"ATen fp16     : {t4:8.3f} ms  {flop / t4 / 1e9:8.1f} TFLOP/s (unfused cuBLAS fp16 chain)")
