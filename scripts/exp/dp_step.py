"""Experiment (torchrun, N ranks): time of the optimiser leg alone — fused peer-memory kernel vs NCCL all-reduce + dense Adam."""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, ".")
from nerfpp_b200 import ops, parallel
from nerfpp_b200.pipeline import HashNeRF

rank, world, local = parallel.init_from_env("nccl")
dev = torch.device("cuda", local); torch.cuda.set_device(dev)
a = HashNeRF(device=dev); b = HashNeRF(device=dev)
parallel.PeerShardedOptimizer(b, rank, world)

def timeit(f, n=30):
    for _ in range(5): f()
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3

def nccl_step():
    a.optimizer_step(grad_scale=parallel.allreduce_gradients(a.grads, world))
def nccl_only():
    parallel.allreduce_gradients(a.grads, world)
t_f = timeit(b.optimizer_step_sharded); t_n = timeit(nccl_step); t_ar = timeit(nccl_only)
st = b.peer.flags[24:29].cpu().tolist()
ph = [(st[i + 1] - st[i]) % (1 << 32) / 1e3 for i in range(4)]
print(f"rank {rank} last fused kernel phases us: entry barrier {ph[0]:.1f} | shard loop (CTA 0) {ph[1]:.1f} | exit barrier {ph[2]:.1f} | grad clear {ph[3]:.1f}", flush=True)
if rank == 0:
    print(f"world {world}: fused step {t_f:.1f} us | nccl all-reduce + adam + pack {t_n:.1f} us | all-reduce alone {t_ar:.1f} us | timeout flag {b.flags_timeout()}")
dist.barrier(); dist.destroy_process_group()
