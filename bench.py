#!/usr/bin/env python
"""Headline benchmark: HashNeRF training throughput (rays/s) on the ray-batch hot path, BASELINE config C2/C3.

    python bench.py [--gpus N] [--steps K] [--warmup W]            # this repo's sm_100a path
    python bench.py --impl reference [--steps K] [--warmup W]      # the reference's own LibTorch CPU path (oracle/_ref)

One step = render(4096 rays: 64 coarse + 128 importance samples, two network passes) + huber + backward + all-reduce
(N>1) + Adam, i.e. NeRFExecutor::Train's loop body (reference src/NeRFExecutor.h:868-996) in the parity configuration.
Prints ONE JSON line (rank 0).  See DESIGN.md §measurement for the definitions of every field.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
for p in (str(ROOT), str(ROOT / "oracle"), str(ROOT / "oracle" / "_ref")):
    if p not in sys.path:
        sys.path.insert(0, p)

RAYS_PER_GPU = 4096
N_SAMPLES, N_IMPORTANCE = 64, 128
BBOX = (-1.5, -1.5, -1.5, 1.5, 1.5, 1.5)
CPU_SAMPLE_RAYS = 256   # bounded sample of the 4096-ray step for the CPU arms

# dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed `ncu --set full` captures (profiles/), largest launch
NCU_TRAFFIC_BYTES = {
    "hash_encode_fwd": 19.59e6 + 3.89e6,    # profiles/r1_hash_fwd_v3_ncu.txt, 786 432-point launch
    "hash_encode_bwd": 89.12e6 + 2.39e6,    # profiles/r1_hash_bwd_v2_ncu.txt
}
# algorithmic bytes per unit (SURVEY §8d with this repo's fp16 encoding output)
BYTES_PER_POINT = {
    "hash_encode_fwd": 12 + 512 + 64 + 1,   # xyz in, 16 lvl x 8 corners x 2 feat x fp16 gathered, fp16 [32] out, keep byte
    "hash_encode_bwd": 12 + 64 + 2 * 512,   # xyz in, bf16 [32] grad in, read-modify-write of the 8x16 gathered entries
}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed regions (B200_PROFILING.md recipe): one `nvidia-smi -lms` process
    streaming a sample every 20 ms, started before the first timed region and stopped after the last."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.rows = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index), "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:  # noqa: BLE001 - sampling must never break the bench
            self.proc = None

    def stop(self) -> dict:
        if self.proc is not None:
            try:
                self.proc.terminate()
                out, _ = self.proc.communicate(timeout=10)
                self.rows = [[c.strip() for c in line.split(",")] for line in out.splitlines() if line.count(",") >= 6]
            except Exception:  # noqa: BLE001
                self.proc.kill()
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm = [float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[3 + i].lower().startswith("active") for r in self.rows if len(r) > 3 + i)]
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": float(self.rows[0][1]), "samples": len(self.rows),
                "power_w_max": max(float(r[2]) for r in self.rows), "reasons": reasons,
                "window": "resident + roofline + e2e timed regions"}


def reference_cpu_rays_per_s(steps: int, warmup: int, rays: int = CPU_SAMPLE_RAYS):
    """The reference's own implementation of the step — NeRFRenderer<HashEmbedder,SHEncoder,NeRFSmall> + huber + backward +
    torch::optim::Adam, LibTorch CPU fp32 on all host threads — on a bounded `rays`-ray sample."""
    import torch
    import nerfpp_ref_cpu as R
    R.set_num_threads(os.cpu_count() or 1)
    R.manual_seed(42)
    devnull = os.open(os.devnull, os.O_WRONLY)   # the reference prints parameter names from Trainable::Initialize
    saved = os.dup(1)
    os.dup2(devnull, 1)
    try:
        pipe = R.make_hash_cpu(torch.tensor(BBOX), 16, 2, 19, 16, 512, 4, 2, 64, 15, 3, 64)
        pipe.init_model()
    finally:
        os.dup2(saved, 1)
        os.close(devnull)
    from nerfpp_b200.pipeline import synthetic_rays
    o, d, tgt = synthetic_rays(rays, device="cpu", seed=0)
    if warmup:
        pipe.train_steps(o, d, tgt, warmup, N_SAMPLES, N_IMPORTANCE, 4096, True, 1e-2, 250)
    secs, _ = pipe.train_steps(o, d, tgt, steps, N_SAMPLES, N_IMPORTANCE, 4096, True, 1e-2, 250)
    total = sum(secs)
    return rays * steps / total, total / steps, R.get_num_threads()


def run_reference(args) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    try:
        value, sec_per_step, threads = reference_cpu_rays_per_s(args.steps, args.warmup)
    except ImportError as e:
        print(json.dumps({"impl": "reference", "unavailable": f"oracle/_ref/nerfpp_ref_cpu.so not importable: {e}"}))
        return
    sample = f"{CPU_SAMPLE_RAYS}-ray sample of the {RAYS_PER_GPU}-ray step, {args.steps} steps after {args.warmup} warm-up"
    print(json.dumps({
        "impl": "reference", "metric": "train_rays_per_s", "value": value, "unit": "rays/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": sec_per_step * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": "HashNeRF train step (L16 T2^19 F2 hash grid, SH4, 2x64 + 3x64 MLP), 64+128 samples/ray, "
                               f"{CPU_SAMPLE_RAYS}-ray bounded sample of the 4096-ray batch, LibTorch CPU"},
        "cpu_baseline": {"value": value, "unit": "rays/s", "cores": threads, "kind": "reference", "sample": sample},
        "e2e": {"value": value, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def _graph_timed(fn, reps: int) -> float:
    """ms per launch of `fn` (C-ABI launches on torch's current stream): `reps` launches captured in ONE CUDA graph and replayed, so
    the figure is device time of back-to-back launches, not ctypes / allocator time on the host (these kernels take 5-400 us)."""
    import torch
    fn()
    torch.cuda.synchronize()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        fn()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=side):
            for _ in range(reps):
                fn()
    torch.cuda.current_stream().wait_stream(side)
    graph.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    graph.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def render_ops_leg(hbm_peak: float, reps: int = 20) -> dict:
    """HBM roofline of the CustomOps / RayUtils rendering kernels (compositing forward / backward, sample_pdf + merge), each timed alone
    with a CUDA-event pair over `reps` launches at two sizes: the fine pass of one 4096-ray training step (64+128 samples) and a
    262 144-ray render tile (BASELINE C4: one GPU's share of a 1080p frame, 64+64 samples).  Algorithmic bytes per SURVEY §8(d):
    forward 24 B/sample + 44 B/ray, backward 40 B/sample + 12 B/ray of map gradients, sampler 4 B x (2S + N + S+N) per ray.
    The same buffers are re-read by every launch: the small size (19 MB) is L2-resident, as it is inside the training step; the
    large size (0.8-1.3 GB per launch) cannot be."""
    import torch
    from nerfpp_b200 import ops

    def timed(fn):
        return _graph_timed(fn, reps)

    out = {"unit": "GB/s", "peak": hbm_peak, "sizes": {}}
    g = torch.Generator().manual_seed(0)
    for tag, rays, s_c, n_imp in (("train_fine_4096x192", 4096, 64, 128), ("render_tile_262144x128", 262144, 64, 64)):
        s = s_c + n_imp
        raw = torch.randn(rays, s, 4, generator=g).cuda()
        z = (2 + torch.sort(torch.rand(rays, s, generator=g) * 4, -1).values).cuda()
        d = torch.randn(rays, 3, generator=g).cuda()
        g_rgb = torch.randn(rays, 3, generator=g).cuda()
        d_raw = torch.empty_like(raw)
        zc = z[:, :s_c].contiguous()
        w = torch.rand(rays, s_c, generator=g).cuda()
        u = torch.linspace(0, 1, n_imp).cuda()
        from nerfpp_b200.cabi import lib, ptr, stream
        maps = [torch.empty(rays, 3, device="cuda")] + [torch.empty(rays, device="cuda") for _ in range(3)] + [torch.empty(rays, s, device="cuda")]
        merged = torch.empty(rays, s, device="cuda")
        ms_f = timed(lambda: lib().nrf_composite_fwd(ptr(raw), 4, ptr(z), ptr(d), None, 0.0, 0, rays, s, *[ptr(m) for m in maps], stream()))
        ms_b = timed(lambda: lib().nrf_composite_bwd(ptr(raw), 4, ptr(z), ptr(d), None, 0.0, 0, rays, s, ptr(g_rgb), None, None, None, None, ptr(d_raw),
                                                       stream()))
        ms_s = timed(lambda: lib().nrf_sample_pdf_merge(ptr(zc), ptr(w), ptr(u), 0, rays, s_c, n_imp, None, ptr(merged), stream()))
        by_f, by_b, by_s = rays * (s * 24 + 44), rays * (s * 40 + 24), rays * 4 * (2 * s_c + n_imp + s)
        out["sizes"][tag] = {
            "composite_fwd": {"ms": ms_f, "achieved": by_f / ms_f / 1e6, "frac": by_f / ms_f / 1e6 / hbm_peak, "bytes": by_f},
            "composite_bwd": {"ms": ms_b, "achieved": by_b / ms_b / 1e6, "frac": by_b / ms_b / 1e6 / hbm_peak, "bytes": by_b},
            "sample_pdf_merge": {"ms": ms_s, "achieved": by_s / ms_s / 1e6, "frac": by_s / ms_s / 1e6 / hbm_peak, "bytes": by_s}}
        del raw, z, d_raw
    return out


def classic_nerf_leg(tf_peak: float, rows: int = 1024 * 192, reps: int = 10) -> dict:
    """BASELINE C1's network (NeRFImpl 8x256, src/NeRF.cpp:92-126) at one 1024-ray training batch of 64+128 samples (196 608 rows):
    inference forward, training forward (stores the layer inputs) and backward (gradient chain + weight gradients), each timed with
    a CUDA-event pair over `reps` launches.  1.187 MFLOP/row forward, 2x that backward (no input gradient: layer 0 has dW only)."""
    import ctypes
    import math

    import torch
    from nerfpp_b200 import ops
    g = torch.Generator().manual_seed(0)
    shapes = {f"model_pts_linears_{i}": (256, 63 if i == 0 else (319 if i == 5 else 256)) for i in range(8)}
    shapes.update({"model_feature_linear": (256, 256), "model_alpha_linear": (1, 256), "model_views_linears_0": (128, 283), "model_rgb_linear": (3, 128)})
    p = {}
    for name, (o, i) in shapes.items():
        p[name + ".weight"] = (torch.randn(o, i, generator=g) * math.sqrt(2.0 / i)).cuda()
        p[name + ".bias"] = (torch.randn(o, generator=g) * 0.1).cuda()
    x = (torch.rand(rows, 90, generator=g) * 2 - 1).cuda()
    gout = (torch.randn(rows, 4, generator=g) * 1e-3).cuda()
    packed, packed_t = ops.mlp_nerf_pack(p), ops.mlp_nerf_pack(p, train=True)
    grads = {k: torch.zeros_like(v) for k, v in p.items()}
    out, saved = ops.mlp_nerf_fwd_train(packed_t, x)
    ws = torch.empty(ops.lib().nrf_mlp_nerf_bwd_workspace_bytes(ctypes.byref(ops.mlp_nerf_shape()), rows), dtype=torch.uint8, device="cuda")

    def timed(fn):
        fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    ms_inf = timed(lambda: ops.mlp_nerf_fwd(packed, x, out=out))
    ms_fwd = timed(lambda: ops.mlp_nerf_fwd_train(packed_t, x))
    ms_bwd = timed(lambda: ops.mlp_nerf_bwd(packed_t, saved, gout, grads, workspace=ws))
    fl = rows * 1.186816e6
    bwd_flop = 2 * fl - rows * 2 * (63 * 256 + 63 * 256 + 27 * 128)      # no dX for the embedded inputs
    return {"rows": rows, "config": "C1 network: 8x256 + skip + view branch, one 1024-ray batch x (64+128) samples",
            "fwd_inference": {"ms": ms_inf, "achieved": fl / ms_inf / 1e9, "frac": fl / ms_inf / 1e9 / tf_peak, "dtype": "fp16"},
            "fwd_train": {"ms": ms_fwd, "achieved": fl / ms_fwd / 1e9, "frac": fl / ms_fwd / 1e9 / tf_peak, "dtype": "bf16",
                          "saved_bytes": int(saved.numel())},
            "bwd": {"ms": ms_bwd, "achieved": bwd_flop / ms_bwd / 1e9, "frac": bwd_flop / ms_bwd / 1e9 / tf_peak, "dtype": "bf16",
                    "kernels": "mlp_nerf_bwd_chain_kernel + mlp_nerf_bwd_dw_kernel"}}


def lerf_leg(tf_peak: float, rays: int = 1024, reps: int = 10) -> dict:
    """BASELINE C5's language head (LeRF(32, 2, 256, 512, 128), src/LeRF.cpp:28-111) at one 1024-ray batch of 64 + 128 samples: the three
    stage programs of the fused tcgen05 kernel and the per-ray finish, each timed with a CUDA-event pair over `reps` launches, and the whole
    LeRFRenderer::RenderRays (inference).  FLOP counts are the REFERENCE's (the full 512-wide last layer), not the issued ones."""
    import torch
    from nerfpp_b200 import ops
    from nerfpp_b200.lerf import LeRFField
    f = LeRFField(seed=0)
    g = torch.Generator().manual_seed(0)
    for v in f.weights.values():
        v.copy_((torch.randn(v.shape, generator=g) * (2.0 / v.shape[1]) ** 0.5).cuda())
    f.table.copy_((torch.rand(f.n_table, generator=g) * 2 - 1).cuda())
    f.refresh()
    n_c, n_f = rays * 64, rays * 192
    enc_c, enc_f = torch.randn(n_c, 128, generator=g).half().cuda(), torch.randn(n_f, 128, generator=g).half().cuda()
    w = torch.rand(rays, 192, generator=g).cuda()

    def timed(fn):
        fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    ms_sigma = timed(lambda: ops.lerf_sigma_fwd(f.packed, enc_c))
    ms_hidden = timed(lambda: ops.lerf_hidden_fwd(f.packed, enc_f))
    _, hidden, q = ops.lerf_hidden_fwd(f.packed, enc_f)
    ms_finish = timed(lambda: ops.lerf_render_embedding(f.packed, w, hidden, q))
    ms_raw = timed(lambda: ops.lerf_fwd(f.packed, enc_f))
    from nerfpp_b200.pipeline import synthetic_rays
    o, d, _ = synthetic_rays(rays, seed=3)
    ms_render = timed(lambda: f.render_rays(o, d, return_weights=False))
    # per-kernel shares of the chunk: one CUDA-event pair around every C-ABI call over `reps` eager chunks
    kt = ops.KernelTimer()
    ops.set_timer(kt)
    for _ in range(reps):
        f.render_rays(o, d, return_weights=False)
    torch.cuda.synchronize()
    ops.set_timer(None)
    kernels_ms = {k: round(ms / reps, 4) for k, (_, ms) in sorted(kt.summary().items(), key=lambda kv: -kv[1][1])}
    f.capture_render(rays)
    ms_render_graph = timed(lambda: f.render_rays_graph(o, d))
    fl_sigma, fl_full = 2 * (128 * 256 + 256 * 33), 2 * (128 * 256 + 256 * 33 + 160 * 256 + 256 * 512)

    def tf(rows, flop, ms):
        return rows * flop / ms / 1e9

    return {"rays": rays, "config": "C5 language head: LeRF(32,2,256,512,128) on a 16x8 hash grid, 1024 rays x (64 coarse + 192 fine) samples, inference",
            "sigma_program": {"rows": n_c, "ms": ms_sigma, "achieved": tf(n_c, fl_sigma, ms_sigma), "frac": tf(n_c, fl_sigma, ms_sigma) / tf_peak},
            "hidden_program_plus_finish": {"rows": n_f, "ms": ms_hidden + ms_finish, "ms_tc_kernel": ms_hidden, "ms_per_ray_finish": ms_finish,
                                           "achieved": tf(n_f, fl_full, ms_hidden + ms_finish), "frac": tf(n_f, fl_full, ms_hidden + ms_finish) / tf_peak,
                                           "note": "reference-equivalent FLOPs (512-wide last layer); issued: 303 kFLOP/row (G = W^T W norm layer)"},
            "raw_program": {"rows": n_f, "ms": ms_raw, "achieved": tf(n_f, fl_full, ms_raw), "frac": tf(n_f, fl_full, ms_raw) / tf_peak,
                            "hbm_gbs": n_f * (513 * 4 + 256) / ms_raw / 1e6, "note": "writes raw_le [N,513] fp32 (the reference's layout)"},
            "render_rays": {"ms": ms_render_graph, "rays_per_s": rays / ms_render_graph * 1e3, "ms_eager": ms_render, "kernels_ms_per_chunk": kernels_ms, "dtype": "fp16 operands, fp32 accumulate",
                            "note": "LeRFRenderer::RenderRays of one 1024-ray chunk (11 kernels) replayed as one CUDA graph; ms_eager = the same through 11 ctypes calls"}}


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--rays", type=int, default=RAYS_PER_GPU, help="rays per GPU per step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="launch every kernel eagerly instead of replaying the captured step")
    ap.add_argument("--dp", default="fused", choices=["fused", "nccl"],
                    help="N>1 optimiser step: one peer-memory kernel (reduce-scatter + Adam + shadow all-gather) or NCCL all-reduce + dense Adam")
    args = ap.parse_args()
    # stdout carries exactly ONE JSON line: everything native libraries print on file descriptor 1 while the job runs (NCCL's version
    # banner, for one) is sent to stderr; the line itself goes to the saved descriptor at the end
    sys.stdout.flush()
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    sys.stdout = real_stdout
    if args.impl == "reference":
        run_reference(args)
        return

    import torch
    from nerfpp_b200 import cabi, ops, parallel
    from nerfpp_b200.pipeline import HashNeRF, synthetic_rays

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback (use --impl reference for the CPU arm)")
    rank, world, local_rank = parallel.init_from_env()
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    warmup = max(args.warmup, 3)
    R = args.rays

    model = HashNeRF(BBOX, n_samples=N_SAMPLES, n_importance=N_IMPORTANCE, device=dev, seed=42)
    parallel.broadcast_parameters(model.params, world)
    model.refresh()

    dp_mode = "single"
    if world > 1:
        dp_mode = "nccl all-reduce of the flat gradient + dense Adam on every rank"
        if args.dp == "fused":
            why = ""
            try:
                peer = parallel.PeerShardedOptimizer(model, rank, world)
            except Exception as e:  # noqa: BLE001 - symmetric memory unavailable on this rank
                peer, why = None, f"{type(e).__name__}: {e}"
            have = torch.tensor([int(peer is not None)], dtype=torch.int32, device=dev)
            torch.distributed.all_reduce(have, op=torch.distributed.ReduceOp.MIN)      # all ranks or none
            if bool(have.item()) and peer.self_test(model):
                dp_mode = "fused peer-memory kernel per rank: reduce-scatter(grad) + Adam(1/N shard) + all-gather(fp16 shadow) over NVLink, no NCCL in the step"
            else:
                model.peer = None        # stay on the NCCL path (the symmetric buffers remain ordinary device memory) and say so
                dp_mode += f" (fused path unavailable: {why or 'peer self-test failed on some rank'})"

    pool = 8  # distinct pre-generated ray batches per rank, cycled
    dev_batches = [synthetic_rays(R, device=dev, seed=1000 * rank + i) for i in range(pool)]
    host_batches = [tuple(t.cpu().pin_memory() for t in b) for b in dev_batches]
    stage = tuple(torch.empty_like(t) for t in dev_batches[0])

    def eager_step(batch):
        model.forward_backward(*batch)
        if model.peer is not None:
            model.optimizer_step_sharded()
        else:
            scale = parallel.allreduce_gradients(model.grads, world)
            model.optimizer_step(grad_scale=scale)

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            torch.distributed.barrier()

    # ---- warm-up, instrumented: every kernel gets an event pair so the dominant one can be named
    timer = ops.KernelTimer()
    ops.set_timer(timer)
    for i in range(warmup):
        eager_step(dev_batches[i % pool])
    torch.cuda.synchronize()
    ops.set_timer(None)
    breakdown = {k: {"launches_per_step": n / warmup, "ms_per_step": ms / warmup} for k, (n, ms) in timer.summary().items()}
    dominant = max(("hash_encode_fwd", "hash_encode_bwd"), key=lambda k: breakdown.get(k, {"ms_per_step": 0})["ms_per_step"])

    # ---- the step as the public API runs it: one CUDA-graph replay (two around the all-reduce when world > 1)
    use_graph = not args.no_graph
    if use_graph:
        model.capture_train_step(R, world, lambda g: parallel.allreduce_gradients(g, world))
        step = lambda batch: model.train_step_graph(*batch)   # noqa: E731
        for i in range(3):
            step(dev_batches[i % pool])
        launches_per_step = model.graph_kernels_per_step
    else:
        step = eager_step
        l0 = cabi.launch_count()
        step(dev_batches[0])
        launches_per_step = cabi.launch_count() - l0 + 1

    # ---- timed region A: K steps, inputs resident in HBM
    sampler = ClockSampler(local_rank)
    sync_all()
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        step(dev_batches[i % pool])
    e1.record()
    sync_all()
    ms_total = parallel.max_over_ranks(e0.elapsed_time(e1), world, dev)
    loss_resident = float(model.loss)

    # ---- roofline leg: the same steps launched eagerly so the dominant kernel can carry a CUDA-event pair on its stream
    # (a graph replay cannot); same buffers, same sizes, run back to back with the timed region
    roof_steps = min(args.steps, 50)
    timer = ops.KernelTimer(only=[dominant, "mlp_small_fwd", "mlp_small_bwd"])
    ops.set_timer(timer)
    sync_all()
    e4, e5 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e4.record()
    for i in range(roof_steps):
        eager_step(dev_batches[i % pool])
    e5.record()
    sync_all()
    ops.set_timer(None)
    ms_eager = e4.elapsed_time(e5) / roof_steps
    n_launch, ms_kernel = timer.summary()[dominant]
    _, ms_mlp_fwd = timer.summary()["mlp_small_fwd"]
    _, ms_mlp_bwd = timer.summary()["mlp_small_bwd"]

    # ---- timed region B (e2e): the public API with HOST buffers — pinned H2D of the step's rays/targets and a D2H read
    # of the loss inside the timed region, every step
    sync_all()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record()
    loss_host = 0.0
    for i in range(args.steps):
        hb = host_batches[i % pool]
        if use_graph:
            step(hb)                    # train_step_graph copies the pinned host rays/targets into its static inputs
        else:
            for dst, src in zip(stage, hb):
                dst.copy_(src, non_blocking=True)
            step(stage)
        loss_host = float(model.loss)   # device -> host read of the step's result
    e3.record()
    sync_all()
    clocks = sampler.stop()
    ms_e2e = parallel.max_over_ranks(e2.elapsed_time(e3), world, dev)
    launches = launches_per_step * args.steps   # kernels inside timed region A (graph: kernel nodes per replay x replays)

    # ---- render leg (BASELINE C4): one 1920x1080 frame, image rows sharded over the ranks, 64 coarse + 64 importance samples
    # (the fine pass evaluates 128 samples/ray), no communication until the final gather of the maps to rank 0
    import math
    H, W = 1080, 1920
    focal = 0.5 * W / math.tan(0.5 * 0.6911)
    K = [[focal, 0, 0.5 * W], [0, focal, 0.5 * H], [0, 0, 1]]
    c2w = torch.eye(4)
    c2w[2, 3] = 4.0
    r0, r1 = parallel.shard_bounds(H, rank, world)
    frames = 3

    def render_frame():
        maps = model.render_image(H, W, K, c2w, chunk=131072, row_begin=r0, row_end=r1, n_importance=64)
        return parallel.gather_rows(maps["rgb"], H * W, rank, world, unit=W) if world > 1 else maps["rgb"]

    render_frame()
    sync_all()
    e6, e7 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e6.record()
    for _ in range(frames):
        img = render_frame()
    e7.record()
    sync_all()
    ms_render = parallel.max_over_ranks(e6.elapsed_time(e7), world, dev) / frames
    render = {"metric": "render_msamples_per_s", "value": H * W * (N_SAMPLES + N_SAMPLES + 64) / (ms_render * 1e3), "unit": "Msamples/s",
              "frames_per_s": 1e3 / ms_render, "ms_per_frame": ms_render, "frames": frames,
              "config": "1920x1080 frame, image rows sharded over the GPUs, 64 coarse + 64 importance samples (192 network evaluations/ray), "
                        "131072-ray chunks, rgb gathered on rank 0; untrained (random-init) model"}

    # ---- LeRF render leg (BASELINE C5's field, the C4 frame): RenderedLangEmbedding of one 1920x1080 frame, image rows sharded over the ranks,
    # 64 coarse + 128 importance samples (256 head evaluations per ray), no communication until the final gather of the [rows, 512] map
    render_lerf = None
    try:
        from nerfpp_b200.lerf import LeRFField
        field = LeRFField(BBOX, seed=0)                    # same seed on every rank: identical table, primes and weights without a broadcast
        g = torch.Generator().manual_seed(0)
        for v in field.weights.values():
            v.copy_((torch.randn(v.shape, generator=g) * (2.0 / v.shape[1]) ** 0.5).to(dev))
        field.table.copy_((torch.rand(field.n_table, generator=g) * 2 - 1).to(dev))
        field.refresh()
        field.render_image(H, W, K, c2w, chunk=8192, row_begin=r0, row_end=min(r0 + 8, r1))      # warm-up on a few rows
        torch.cuda.synchronize(dev)
        ready = True
    except Exception as e:  # noqa: BLE001 - a secondary leg must not take the headline line down
        ready, render_lerf = False, {"error": f"{type(e).__name__}: {e}"}
    if parallel.all_ranks_ready(ready, world, dev):                    # else all ranks skip the leg together

        def lerf_frame():
            maps = field.render_image(H, W, K, c2w, chunk=8192, row_begin=r0, row_end=r1)
            return parallel.gather_rows(maps["rendered"], H * W, rank, world, unit=W) if world > 1 else maps["rendered"]

        sync_all()
        e8, e9 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e8.record()
        emb = lerf_frame()
        e9.record()
        sync_all()
        ms_lerf = parallel.max_over_ranks(e8.elapsed_time(e9), world, dev)
        render_lerf = {"metric": "lerf_render_rays_per_s", "value": H * W / (ms_lerf * 1e-3), "unit": "rays/s", "ms_per_frame": ms_lerf, "frames": 1,
                       "msamples_per_s": H * W * (N_SAMPLES + N_SAMPLES + N_IMPORTANCE) / (ms_lerf * 1e3),
                       "config": "1920x1080 frame of RenderedLangEmbedding [H,W,512] fp32, LeRF(32,2,256,512,128) on a 16x8 T2^19 language grid, image rows sharded "
                                 "over the GPUs, 64 coarse (density-only head) + 192 fine samples per ray, 8192-ray chunks, map gathered on rank 0; random-init field"}
        del emb

    if rank != 0:
        return

    peaks_path = ROOT / "MEASURED_PEAKS.json"
    if peaks_path.exists():
        peak_gbs, peak_src = json.loads(peaks_path.read_text())["hbm_gbs"], "MEASURED_PEAKS.json hbm_gbs (sustained copy)"
    else:
        peak_gbs, peak_src = 6650.0, "fallback B200_PROFILING.md"
    pts_per_step = R * (N_SAMPLES + N_SAMPLES + N_IMPORTANCE) if dominant == "hash_encode_fwd" else R * (N_SAMPLES + N_IMPORTANCE)
    bytes_per_step = BYTES_PER_POINT[dominant] * pts_per_step
    kl_per_step = n_launch / roof_steps
    achieved = (bytes_per_step * roof_steps / 1e9) / (ms_kernel / 1e3)
    roofline = {"bound": "hbm", "kernel": dominant, "achieved": achieved, "peak": peak_gbs, "unit": "GB/s", "frac": achieved / peak_gbs,
                "traffic": NCU_TRAFFIC_BYTES.get(dominant), "peak_source": peak_src, "bytes_per_launch": bytes_per_step / kl_per_step,
                "ms_per_launch": ms_kernel / n_launch, "share_of_step": (ms_kernel / roof_steps) / (ms_total / args.steps),
                "timing": f"CUDA-event pair around every launch over {roof_steps} eagerly launched steps run right after the timed region "
                          f"({ms_eager:.3f} ms/step eager)",
                "l2_note": "the table shadow (17 MiB) is L2-resident: the kernel is bound by L1 tag lookups (forward) / L2 RED operations (backward), "
                           "not by HBM. Ceilings measured on B200 with scripts/exp/{gather,red}_bench.cu: 280 G random 4 B gathers/s, 209 G REDs/s "
                           "(any operand width); hash_fwd issues 301 G sector lookups/s, hash_bwd 206 G REDs/s (DESIGN.md §4)"}

    # the tensor-core side of the step: fused NeRFSmall forward (tcgen05) and backward (mma.sync), useful FLOPs only
    # (18 688 FLOP/point forward over coarse + fine points, 37 376 FLOP/point backward over the fine points; the backward also
    # recomputes the forward, which is not counted)
    peaks = json.loads(peaks_path.read_text()) if peaks_path.exists() else {}
    tf_peak = peaks.get("bf16_tflops", 1590.0)
    pts_fwd, pts_bwd = R * (2 * N_SAMPLES + N_IMPORTANCE), R * (N_SAMPLES + N_IMPORTANCE)
    tf_fwd = pts_fwd * 18688 * roof_steps / (ms_mlp_fwd * 1e-3) / 1e12
    tf_bwd = pts_bwd * 37376 * roof_steps / (ms_mlp_bwd * 1e-3) / 1e12
    roofline_tensor = {"bound": "tensor", "unit": "TFLOP/s", "peak": tf_peak,
                       "peak_source": "MEASURED_PEAKS.json bf16_tflops (cuBLAS burst)" if peaks else "fallback B200_PROFILING.md",
                       "mlp_small_fwd": {"achieved": tf_fwd, "frac": tf_fwd / tf_peak, "ms_per_step": ms_mlp_fwd / roof_steps, "pipe": "tcgen05"},
                       "mlp_small_bwd": {"achieved": tf_bwd, "frac": tf_bwd / tf_peak, "ms_per_step": ms_mlp_bwd / roof_steps, "pipe": "mma.sync (HMMA)"},
                       "note": "event pairs include both launches of the forward (coarse + fine) and their launch gaps; ncu per-launch figures in profiles/"}
    # what actually bounds the tcgen05 forward: its epilogues read every layer's fp32 accumulators out of TMEM at 64 B/clk/SM
    # (B300_MICROARCH.md, LDTM throughput): 224 accumulator columns x 4 B per point
    sm_mhz = peaks.get("sm_max_mhz", 1965.0)
    tmem_floor_ms = pts_fwd * 224 * 4 / (148 * 64 * sm_mhz * 1e6) * 1e3
    roofline_tensor["mlp_small_fwd"]["tmem_read_roofline"] = {"bytes_per_point": 896, "peak": "64 B/clk/SM x 148 SMs", "floor_ms_per_step": tmem_floor_ms,
                                                               "frac": tmem_floor_ms / (ms_mlp_fwd / roof_steps)}

    roofline_render_ops = None
    if rank == 0:
        roofline_tensor["mlp_nerf"] = classic_nerf_leg(tf_peak)
        try:
            roofline_tensor["lerf_head"] = lerf_leg(tf_peak)
        except Exception as e:  # noqa: BLE001 - a secondary leg must not take the headline line down
            roofline_tensor["lerf_head"] = {"error": f"{type(e).__name__}: {e}"}
        roofline_render_ops = render_ops_leg(roofline["peak"])

    cpu_baseline = None
    if world == 1 and not args.no_cpu_baseline:
        try:
            v, sps, threads = reference_cpu_rays_per_s(steps=6, warmup=1)
            cpu_baseline = {"value": v, "unit": "rays/s", "cores": threads, "kind": "reference",
                            "sample": f"{CPU_SAMPLE_RAYS}-ray sample of the {RAYS_PER_GPU}-ray step, 6 steps after 1 warm-up ({sps:.2f} s/step)"}
        except ImportError as e:
            cpu_baseline = {"value": None, "unit": "rays/s", "cores": 0, "kind": "reference", "sample": f"oracle/_ref not importable: {e}"}

    rays_total = R * world * args.steps
    h2d = sum(t.numel() * t.element_size() for t in host_batches[0])
    print(json.dumps({
        "metric": "train_rays_per_s", "value": rays_total / (ms_total / 1e3), "unit": "rays/s", "n_gpus": world, "steps": args.steps,
        "warmup": warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16", "data": "synthetic",
        "config": {"workload": "C2/C3 HashNeRF training: L16 F2 T2^19 hash grid 16->512 + SH deg 4 + NeRFSmall 32->64->16 | 31->64->64->3, "
                               f"{R} rays/GPU/step of an 800x800 view, 64 coarse + 128 importance samples, huber + Adam(0.9,0.99,1e-15)",
                   "rays_per_gpu": R, "global_rays": R * world, "parallelism": f"ray-sharded dp{world}: {dp_mode}",
                   "l2": "not flushed explicitly: each step streams ~300 MB (Adam pass over 8.9M params + moments + gradient) through the 126 MB L2",
                   "precision": "fp16 hash table reads / encodings, bf16 tensor-core MLP with fp32 accumulate, fp32 everything else"},
        "e2e": {"value": rays_total / (ms_e2e / 1e3), "unit": "rays/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": launches * world, "graph_replay": use_graph, "kernels_per_step": launches_per_step, "roofline": roofline, "roofline_tensor": roofline_tensor, "roofline_render_ops": roofline_render_ops, "cpu_baseline": cpu_baseline, "clocks": clocks,
        "kernels_ms_per_step": {k: round(v["ms_per_step"], 4) for k, v in sorted(breakdown.items(), key=lambda kv: -kv[1]["ms_per_step"])},
        "final_loss": {"resident": loss_resident, "e2e": loss_host}, "render": render, "render_lerf": render_lerf,
    }))


if __name__ == "__main__":
    main()
