"""Multi-GPU rollout: trajectories are independent, so the batch is split into contiguous shards, one per
rank (one process per GPU), weights are replicated, and there is NO collective on the inference path
(SURVEY.md §8e).  Gathering the outputs is optional and is the only communication."""
from typing import Optional, Tuple

import torch


def shard_bounds(n_items: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Contiguous, balanced [lo, hi) slice of `n_items` owned by `rank` (first n % world ranks get one extra)."""
    if world_size <= 0 or not (0 <= rank < world_size):
        raise ValueError(f"bad rank/world_size {rank}/{world_size}")
    base, rem = divmod(n_items, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


class ShardedRollout:
    """embed(x0) -> generate -> recover for this rank's shard of a global batch of initial states."""

    def __init__(self, embedding_model, transformer, rank: int = 0, world_size: int = 1):
        self.embedding_model = embedding_model
        self.transformer = transformer
        self.rank, self.world_size = rank, world_size

    def local_slice(self, n_global: int) -> slice:
        lo, hi = shard_bounds(n_global, self.world_size, self.rank)
        return slice(lo, hi)

    @torch.no_grad()
    def __call__(self, x0_local: torch.Tensor, horizon: int, use_cache: bool = True, visc: Optional[torch.Tensor] = None,
                 recover: bool = True):
        """x0_local: this rank's initial states.  Returns (embeds [b, horizon+1, E], states or None)."""
        g0 = self.embedding_model.embed(x0_local) if visc is None else self.embedding_model.embed(x0_local, visc)
        out = self.transformer.generate(g0.unsqueeze(1), max_length=horizon + 1, use_cache=use_cache,
                                        return_presents=False)[0]
        states = None
        if recover:
            b, t, e = out.shape
            states = self.embedding_model.recover(out.reshape(b * t, e))
            states = states.view(b, t, *states.shape[1:])
        return out, states

    def gather(self, local: torch.Tensor, n_global: int) -> torch.Tensor:
        """All-gather shards of unequal size along dim 0 (only if the caller wants one tensor)."""
        import torch.distributed as dist
        if self.world_size == 1:
            return local
        sizes = [shard_bounds(n_global, self.world_size, r) for r in range(self.world_size)]
        maxn = max(hi - lo for lo, hi in sizes)
        pad = torch.zeros((maxn,) + tuple(local.shape[1:]), device=local.device, dtype=local.dtype)
        pad[: local.size(0)] = local
        bufs = [torch.empty_like(pad) for _ in range(self.world_size)]
        dist.all_gather(bufs, pad)
        return torch.cat([b[: hi - lo] for b, (lo, hi) in zip(bufs, sizes)], dim=0)


def average_flat_grad(flat_grad: torch.Tensor, world_size: int) -> torch.Tensor:
    """Data-parallel Koopman training step: ONE all-reduce (SUM) of the flat gradient buffer, divided by the
    world size.  With equal shard sizes the mean of the per-shard gradients is the global-batch gradient
    (every loss term is a batch MEAN, and the |K|^2 regulariser is batch independent, so it must be averaged).
    NCCL on GPUs (144,768 bytes: latency bound), gloo in the CPU tests."""
    import torch.distributed as dist
    if world_size > 1:
        dist.all_reduce(flat_grad, op=dist.ReduceOp.SUM)
        flat_grad.div_(world_size)
    return flat_grad
