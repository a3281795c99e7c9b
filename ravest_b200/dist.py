"""Multi-GPU sharding of the sample axis.

Every sample's log-probability depends only on its own theta row and the small read-only
epoch arrays, so rows are partitioned into contiguous blocks, one per rank (one process per
GPU, `torch.distributed`); each rank keeps its own resident copy of the epochs and evaluates
its block with no data-path collective.  The only exchange is ONE all-gather of the S fp64
results (NCCL over NVLink on GPUs, gloo in the CPU tests) so that every rank - and the
host-side sampler - sees all log-probs, as an ensemble half-step needs.

Because a sample's bits never depend on which shard it lands in (csrc/rvlp_kernels.cuh), the
gathered vector is bit-identical for every world size.
"""
from __future__ import annotations

from typing import Callable

ALIGN = 4   # kG in csrc/rvlp_kernels.cuh: shards start on prologue-batch boundaries


def shard_bounds(n_samples: int, world_size: int, rank: int, align: int = ALIGN) -> tuple[int, int]:
    """Contiguous [lo, hi) of rows for `rank`; blocks are multiples of `align` except the last."""
    n_blocks = (n_samples + align - 1) // align
    base, rem = divmod(n_blocks, world_size)
    lo_b = rank * base + min(rank, rem)
    hi_b = lo_b + base + (1 if rank < rem else 0)
    return min(lo_b * align, n_samples), min(hi_b * align, n_samples)


def max_shard(n_samples: int, world_size: int, align: int = ALIGN) -> int:
    return max(shard_bounds(n_samples, world_size, r, align)[1] - shard_bounds(n_samples, world_size, r, align)[0]
               for r in range(world_size))


def sharded_logprob(eval_fn: Callable, theta, n_samples: int | None = None, group=None, theta_is_local: bool = False):
    """Evaluate `eval_fn` on this rank's rows and all-gather the results.

    eval_fn(theta_local[n, ndim]) -> tensor[n] on theta's device.  `theta` is either the full
    [S, ndim] tensor (every rank holds it, e.g. emcee's proposal block) or, with
    theta_is_local=True, just this rank's block of an S = n_samples problem.
    Returns the full [S] tensor on every rank.
    """
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()):
        return eval_fn(theta)
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    S = int(n_samples if n_samples is not None else theta.shape[0])
    lo, hi = shard_bounds(S, world, rank)
    local = theta if theta_is_local else theta[lo:hi]
    if local.shape[0] != hi - lo:
        raise ValueError(f"rank {rank}: local block has {local.shape[0]} rows, expected {hi - lo}")
    part = eval_fn(local) if hi > lo else torch.empty(0, dtype=torch.float64, device=theta.device)
    m = max_shard(S, world)
    if m * world == S:                   # equal shards (the usual case): gather straight into the result
        out = torch.empty(S, dtype=torch.float64, device=part.device)
        dist.all_gather_into_tensor(out, part.contiguous(), group=group)
        return out
    send = torch.full((m,), float("nan"), dtype=torch.float64, device=part.device)
    send[: hi - lo] = part
    recv = torch.empty(world * m, dtype=torch.float64, device=part.device)
    dist.all_gather_into_tensor(recv, send, group=group)
    out = torch.empty(S, dtype=torch.float64, device=part.device)
    for r in range(world):
        a, b = shard_bounds(S, world, r)
        out[a:b] = recv[r * m: r * m + (b - a)]
    return out
