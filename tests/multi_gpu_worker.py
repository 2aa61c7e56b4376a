"""Worker for tests/test_gpu_multi.py (one process per GPU, launched by torch.distributed.run).
Checks the fused data-parallel optimiser (nrf_adam_step_sharded over NVLink peer memory) against the NCCL all-reduce +
dense-Adam path, and the row-sharded render gather."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from nerfpp_b200 import parallel  # noqa: E402
from nerfpp_b200.pipeline import HashNeRF, synthetic_rays  # noqa: E402

BBOX = (-1.5, -1.5, -1.5, 1.5, 1.5, 1.5)


def main():
    rank, world, local = parallel.init_from_env("nccl")
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    R = 512
    batches = [synthetic_rays(R, device=dev, seed=100 * rank + i) for i in range(3)]

    # A: NCCL all-reduce + dense Adam on every rank (eager)
    a = HashNeRF(BBOX, log2_hashmap_size=15, seed=3, device=dev, lrate_decay=1)
    parallel.broadcast_parameters(a.params, world); a.refresh()
    # B: fused peer-memory optimiser (eager), C: the same inside the captured graph
    b = HashNeRF(BBOX, log2_hashmap_size=15, seed=3, device=dev, lrate_decay=1)
    parallel.broadcast_parameters(b.params, world); b.refresh()
    parallel.PeerShardedOptimizer(b, rank, world)
    c = HashNeRF(BBOX, log2_hashmap_size=15, seed=3, device=dev, lrate_decay=1)
    parallel.broadcast_parameters(c.params, world); c.refresh()
    assert parallel.PeerShardedOptimizer(c, rank, world).self_test(c)     # dry step on the zero gradient: parameters untouched
    assert c.step == 0
    c.capture_train_step(R, world)

    la, lb, lc = [], [], []
    for i in range(6):
        batch = batches[i % 3]
        a.forward_backward(*batch)
        a.optimizer_step(grad_scale=parallel.allreduce_gradients(a.grads, world))
        la.append(float(a.loss))
        b.forward_backward(*batch)
        b.optimizer_step_sharded()
        lb.append(float(b.loss))
        lc.append(float(c.train_step_graph(*batch)))
    torch.cuda.synchronize()
    for name, m, losses in (("eager", b, lb), ("graph", c, lc)):
        assert int(m.flags_timeout()) == 0, f"{name}: peer barrier timed out"
        # every rank holds the SAME fp16 shadow, bit for bit
        ref = m.shadow.clone()
        dist.broadcast(ref, src=0)
        assert torch.equal(ref, m.shadow), f"{name}: shadows differ between ranks"
        # the owner's fp32 master agrees with the shadow it published; the MLP tail is replicated
        lo, hi = m.peer.shard_bounds(m.n_table)
        assert torch.equal(m.params[lo:hi].half(), m.shadow[lo:hi]), f"{name}: owner shard / shadow mismatch"
        assert torch.equal(m.params[m.n_table:].half(), m.shadow[m.n_table:])
        tail = m.params[m.n_table:].clone()
        dist.broadcast(tail, src=0)
        assert torch.equal(tail, m.params[m.n_table:]), f"{name}: MLP replicas differ"
        assert float(m.grads.abs().max()) == 0.0, f"{name}: gradient not cleared"
        # against the NCCL path: same training trajectory (Adam is sign-like, so compare the bulk, see test_gpu_pipeline)
        for x, y in zip(losses, la):
            assert abs(x - y) <= 5e-3 * abs(y) + 1e-6, (name, losses, la)
        d = (a.shadow.float() - m.shadow.float()).abs()
        frac = (d > 1e-4).float().mean().item()
        assert frac < 3e-2 and d.median().item() < 1e-6, (name, frac)
        assert m.step == 6 and int(m.sched[0]) == 6

    # row-sharded render + final gather
    H, W = 24, 32
    K = [[40.0, 0, 16.0], [0, 40.0, 12.0], [0, 0, 1]]
    c2w = torch.eye(4); c2w[2, 3] = 4.0
    r0, r1 = parallel.shard_bounds(H, rank, world)
    part = c.render_image(H, W, K, c2w, row_begin=r0, row_end=r1)["rgb"]
    full = parallel.gather_rows(part, H * W, rank, world, unit=W)
    if rank == 0:
        whole = c.render_image(H, W, K, c2w)["rgb"]
        assert torch.equal(full, whole)
        print("MULTI_GPU_WORKER_OK", world, la[-1], lb[-1], lc[-1])
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
