"""Multi-GPU plumbing: one process per GPU, torch.distributed (NCCL over NVLink on the box, gloo in CPU tests).

The reference has no multi-GPU code at all; this is the sharding the hot path admits:
  * training  — rays are independent, so the global ray batch is split contiguously by rank, every rank runs
    forward/backward on its shard, and the ONE exchange step is an all-reduce(sum) of the flat gradient vector
    [reachable hash-table scalars | NeRFSmall weights] (34 MiB + 37 KiB fp32 at the BASELINE shape), scaled by 1/world
    inside the Adam kernel (the loss is a mean over rays, src/NeRFExecutor.h:883-886).  Replicas stay identical
    because they start identical (same seed) and apply the same reduced gradient.
  * rendering — image rows are split into contiguous tiles, no communication until a final gather of the maps.
"""
from __future__ import annotations

import os

import sys

import torch
import torch.distributed as dist


def init_from_env(backend: str | None = None) -> tuple[int, int, int]:
    """(rank, world, local_rank) from the torchrun environment; initialises the default process group if world > 1."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
            dist.init_process_group(backend, device_id=torch.device("cuda", local_rank))
        else:
            dist.init_process_group(backend)
    return rank, world, local_rank


def shard_bounds(n: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous, balanced partition of n units (rays or image rows): the first n % world ranks get one extra."""
    base, extra = divmod(n, world)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def allreduce_gradients(flat_grad: torch.Tensor, world: int) -> float:
    """Sum the flat gradient over ranks in place; returns the scale (1/world) the optimiser must apply."""
    if world > 1:
        dist.all_reduce(flat_grad, op=dist.ReduceOp.SUM)
    return 1.0 / world


def broadcast_parameters(flat_params: torch.Tensor, world: int, src: int = 0) -> None:
    """Rank 0's table / primes / weights become everyone's (the reference draws them at random)."""
    if world > 1:
        dist.broadcast(flat_params, src=src)


def gather_rows(local: torch.Tensor, n_total: int, rank: int, world: int, dst: int = 0, unit: int = 1) -> torch.Tensor | None:
    """Final gather of a row-sharded map ([rows_local, ...]) onto `dst`; shards follow shard_bounds over n_total // unit
    units of `unit` tensor rows each (unit = image width when image rows are the sharded unit)."""
    if world == 1:
        return local
    shapes = [tuple(unit * v for v in shard_bounds(n_total // unit, r, world)) for r in range(world)]
    # equal-size exchange (gloo's gather refuses ragged shards; shard sizes differ by at most one unit): pad to the largest shard, trim on dst
    biggest = max(e - b for b, e in shapes)
    local = local.contiguous()
    if local.shape[0] < biggest:
        local = torch.cat([local, local.new_zeros((biggest - local.shape[0],) + tuple(local.shape[1:]))], 0)
    if rank == dst:
        parts = [torch.empty_like(local) for _ in shapes]
        dist.gather(local, parts, dst=dst)
        return torch.cat([p[:e - b] for p, (b, e) in zip(parts, shapes)], 0)
    dist.gather(local, None, dst=dst)
    return None


def max_over_ranks(value: float, world: int, device) -> float:
    if world == 1:
        return value
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def all_ranks_ready(ok: bool, world: int, device) -> bool:
    """True only if EVERY rank passed ok = True (min over ranks).  Called before the first collective of an optional leg, so that a rank
    whose set-up failed makes all ranks skip it together instead of leaving the others waiting in a gather."""
    return -max_over_ranks(-(1.0 if ok else 0.0), world, device) > 0.5


class PeerShardedOptimizer:
    """Plumbing for nrf_adam_step_sharded: allocates the model's flat gradient, its fp16 shadow and a flag block in
    torch.distributed symmetric memory (peer-mapped over NVLink), exchanges the peer addresses, and re-points the model at
    those buffers.  After this, one kernel per rank performs reduce-scatter(grad) + Adam(owned shard) + all-gather(fp16 shadow);
    NCCL is not on the training path any more (it still does the one-off parameter broadcast and the final render gather)."""

    def __init__(self, model, rank: int, world: int, group=None):
        import torch.distributed._symmetric_memory as symm
        from . import cabi
        group = group or dist.group.WORLD
        dev = model.device
        n = model.params.numel()
        try:                                   # older torch releases need the group enabled explicitly
            symm.enable_symm_mem_for_group(group.group_name)
        except Exception:                      # noqa: BLE001
            pass
        self.grads = symm.empty(n, dtype=torch.float32, device=dev)
        self.shadow = symm.empty(n, dtype=torch.float16, device=dev)
        n_flags = max(int(cabi.lib().nrf_peer_flags_bytes(world)) // 4, 64)
        self.flags = symm.empty(n_flags, dtype=torch.int32, device=dev)
        self.grads.zero_()
        self.flags.zero_()
        self.shadow.copy_(model.shadow)
        handles = [symm.rendezvous(t, group.group_name) for t in (self.grads, self.shadow, self.flags)]
        pg = cabi.PeerGroup()
        pg.world, pg.rank = world, rank
        for field, h, t in zip(("grads", "shadow_f16", "flags"), handles, (self.grads, self.shadow, self.flags)):
            ptrs = list(h.buffer_ptrs)
            off = t.data_ptr() - ptrs[rank]    # offset of the tensor inside this rank's symmetric block (same on every rank)
            assert 0 <= off < (1 << 40), "symmetric-memory tensor is not inside its own block"
            for p in range(world):
                getattr(pg, field)[p] = ptrs[p] + off
        self.pg, self.handles, self.rank, self.world = pg, handles, rank, world
        self.multicast = bool(getattr(handles[0], "has_multicast_support", False))
        model.grads, model.shadow = self.grads, self.shadow
        model.peer = self
        torch.cuda.synchronize(dev)
        dist.barrier(group)

    def self_test(self, model) -> bool:
        """One fused optimiser step on the all-zero gradient before training starts: with zero moments it leaves every parameter
        unchanged but exercises the peer mappings and both barriers.  Returns True when every rank came through (collective)."""
        ok = 1
        try:
            assert model.step == 0 and float(model.grads.abs().max()) == 0.0, "self_test must run before the first step"
            before = model.shadow.clone()
            model.optimizer_step_sharded()
            torch.cuda.synchronize(model.device)
            if model.flags_timeout() or not torch.equal(before, model.shadow):
                print(f"[nerfpp_b200] rank {self.rank}: fused optimiser self-test failed (flags_timeout={model.flags_timeout()}, "
                      f"shadow changed={not torch.equal(before, model.shadow)})", file=sys.stderr, flush=True)
                ok = 0
        except Exception as e:  # noqa: BLE001
            print(f"[nerfpp_b200] rank {self.rank}: fused optimiser self-test raised {type(e).__name__}: {e}", file=sys.stderr, flush=True)
            ok = 0
        model.step = 0
        model._sched_step = -1          # the dry step advanced the device-side counter: re-seed it on the next step
        flag = torch.tensor([ok], dtype=torch.int32, device=model.device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        return bool(flag.item())

    def shard_bounds(self, n_sharded: int) -> tuple[int, int]:
        """[begin, end) scalars of the table this rank owns (same split as the kernel: quads, first ranks one extra)."""
        b, e = shard_bounds(n_sharded // 4, self.rank, self.world)
        return 4 * b, 4 * e
