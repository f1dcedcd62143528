"""K sharded over ranks, one process per GPU: host-staged exchange through any torch.distributed backend.

Each rank owns a contiguous slice of the K samples (global ids offset .. offset + K_local). One tick is three
phases (include/m3p2i_b200.h, m3p2i_phase_*):
    rollout  : local shard -> discounted costs J_local [K_local]
    exchange : all-gather J -> every rank computes identical softmin weights; local weighted partial sums
    finish   : all-reduce(sum) of the packed partial sums [6*T*nu+1] -> identical mean update on every rank
The fast path on GPUs is m3p2i_comm_init + m3p2i_command (NCCL collectives enqueued on the kernel stream); this
class is the transport-agnostic version (gloo on CPU test rigs, or NCCL through torch tensors).
"""
import numpy as np
import torch
import torch.distributed as dist


def shard_bounds(K_global, rank, world):
    if K_global % world:
        raise ValueError(f"num_samples {K_global} must be divisible by the number of ranks {world}")
    Kl = K_global // world
    return Kl, rank * Kl


def attach_peers(backend, group=None):
    """Switch a sharded NativePlanner to the exchange over NVLink peer memory: descriptors travel through
    torch.distributed (any backend), the data path afterwards is stores into peer HBM issued by the kernels
    themselves (include/m3p2i_b200.h, m3p2i_peer_*). One process per GPU. Raises on EVERY rank if any rank failed to map
    its peers (the caller can then fall back to m3p2i_comm_init together)."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    mine = backend.peer_export()
    every = [None] * world
    dist.all_gather_object(every, mine, group=group)
    err = None
    try:
        backend.peer_attach(rank, world, every)
    except Exception as exc:  # e.g. cudaIpc not permitted: all ranks must learn of it, or the next collective hangs
        err = repr(exc)
    errs = [None] * world
    dist.all_gather_object(errs, err, group=group)    # agreement + barrier: no rank starts before all have attached
    bad = [f"rank {r}: {e}" for r, e in enumerate(errs) if e is not None]
    if bad:
        raise RuntimeError("peer-memory exchange unavailable (" + "; ".join(bad) + ")")


class ShardedPlanner:
    def __init__(self, backend, group=None, device="cpu"):
        self.b = backend
        self.group = group
        self.device = device
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        if backend.K * self.world != backend.Kg or backend.cfg.sample_offset != self.rank * backend.K:
            raise ValueError("backend shard does not match this rank: K_local * world == K_global and "
                             "sample_offset == rank * K_local required")

    def command(self, want_cost=True):
        J_local = torch.from_numpy(self.b.phase_rollout()).to(self.device)
        J = torch.empty(self.b.Kg, dtype=torch.float32, device=self.device)
        dist.all_gather_into_tensor(J, J_local, group=self.group) if self.device != "cpu" else \
            dist.all_gather(list(J.view(self.world, -1).unbind(0)), J_local, group=self.group)
        part = torch.from_numpy(self.b.phase_partials(J.cpu().numpy())).to(self.device)
        dist.all_reduce(part, op=dist.ReduceOp.SUM, group=self.group)
        return self.b.phase_finish(part.cpu().numpy())
