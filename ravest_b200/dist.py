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

import os
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


class _RawCuda:
    """A [n] fp64 view of raw device memory for torch.as_tensor (CUDA array interface, zero-copy)."""

    def __init__(self, ptr: int, n: int):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f8", "data": (ptr, False), "version": 2}


class PeerGather:
    """The gathered [S] log-prob vector of every rank, mapped into every rank of `group` by CUDA IPC.

    K1 (`rvlp_logprob_batch_peers`) stores each log-probability of this rank's block straight into every rank's
    vector - 8-byte NVLink peer stores from the kernel's epilogue - and `rvlp_peer_barrier` (one warp, flags in the
    same mapped blocks) tells every rank when all blocks have landed: the all-gather is fused into the kernel, there
    is no collective launch, no staging vector and no second stream.  Two vectors alternate, so a rank that is one step
    ahead never overwrites what a slower rank still reads; a result therefore stays valid until the call after next.
    Construction is collective (every rank of the group, same n_samples); it raises RvlpError when CUDA IPC is not
    available, on every rank alike, and `sharded_logprob` then keeps to the NCCL all-gather.
    """

    FLAG_BYTES = 1024
    TIMEOUT_MS = 120_000    # a rank this late is gone: the barrier raises its flag (timed_out()) instead of hanging

    def __init__(self, n_samples: int, device: int, group=None):
        import ctypes as C
        import torch
        import torch.distributed as dist
        from . import _lib
        self._lib = _lib.load()
        self.S, self.device, self.group = int(n_samples), int(device), group
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        if self.world > 8:
            raise _lib.RvlpError("PeerGather serves up to 8 ranks (one NVSwitch box)")
        self.epoch = 0
        self.base: list[int] = [0] * self.world
        self._opened: list[int] = []
        self._own = 0
        nbytes = 2 * self.S * 8 + self.FLAG_BYTES
        handle = C.create_string_buffer(64)
        ptr = C.c_void_p()
        ok = self._lib.rvlp_peer_alloc(self.device, nbytes, C.byref(ptr), handle) == 0
        if ok:
            self._own = int(ptr.value)
        handles = [None] * self.world
        dist.all_gather_object(handles, (ok, handle.raw), group=group)
        ok = all(h[0] for h in handles)
        if ok:
            for r, (_, raw) in enumerate(handles):
                if r == self.rank:
                    self.base[r] = self._own
                    continue
                p = C.c_void_p()
                if self._lib.rvlp_peer_open(self.device, raw, C.byref(p)) != 0:
                    ok = False
                    break
                self.base[r] = int(p.value)
                self._opened.append(int(p.value))
        flags = [None] * self.world
        dist.all_gather_object(flags, ok, group=group)          # every rank has mapped every block (or nobody goes on)
        if not all(flags):
            self.close()
            raise _lib.RvlpError("CUDA IPC peer mapping is not available on this box: " +
                                 self._lib.rvlp_last_error().decode(errors="replace"))
        self._views = [torch.as_tensor(_RawCuda(self._own + b * self.S * 8, self.S), device=f"cuda:{self.device}")
                       for b in (0, 1)]
        fl = 2 * self.S * 8
        self._flag_ptrs = (C.c_void_p * self.world)(*[b + fl for b in self.base])
        self._outs = [(C.c_void_p * self.world)(*[b + h * self.S * 8 for b in self.base]) for h in (0, 1)]
        self._status = torch.as_tensor(_RawCuda(self._own + fl, self.FLAG_BYTES // 8), device=f"cuda:{self.device}")

    def logprob(self, ctx, theta_local, lo: int):
        """Evaluate this rank's block (rows lo.. of the S) and return the gathered [S] tensor (see the class note)."""
        from ._lib import check, stream_ptr
        self.epoch += 1
        h = self.epoch & 1
        th = ctx._theta(theta_local) if theta_local.shape[0] else theta_local
        if (ctx.k1_variant is None and th.shape[0] >= ctx.AUTOTUNE_MIN_ROWS
                and os.environ.get("RVLP_AUTOTUNE", "1") != "0"):
            ctx.autotune(th)
        st = stream_ptr(self.device)
        check(self._lib.rvlp_logprob_batch_peers(ctx._h, th.data_ptr() if th.shape[0] else None, th.shape[0],
                                                 self._outs[h], self.world, int(lo), st))
        check(self._lib.rvlp_peer_barrier(self.device, self._flag_ptrs, self.world, self.rank, self.epoch,
                                          self.TIMEOUT_MS, st))
        return self._views[h]

    def timed_out(self) -> bool:
        """True when a barrier gave up waiting for a rank (synchronises the device)."""
        import torch
        torch.cuda.synchronize(self.device)
        return bool(self._status.view(torch.int64)[8].item() != 0)

    def close(self):
        for p in self._opened:
            self._lib.rvlp_peer_close(self.device, p)
        self._opened = []
        if self._own:
            self._lib.rvlp_peer_free(self.device, self._own)
            self._own = 0


def _peer_gather_for(ctx, S: int, device: int, group):
    """The PeerGather of (ctx, S, group), built collectively on first use; None when IPC is refused (remembered)."""
    from ._lib import RvlpError
    cache = ctx.__dict__.setdefault("_peer_gathers", {})
    key = (S, id(group))
    if key not in cache:
        try:
            cache[key] = PeerGather(S, device, group)
        except RvlpError:
            cache[key] = None
    return cache[key]


def sharded_logprob(eval_fn: Callable, theta, n_samples: int | None = None, group=None, theta_is_local: bool = False,
                    ctx=None, copy: bool = True, fused: bool | None = None):
    """Evaluate `eval_fn` on this rank's rows and all-gather the results.

    eval_fn(theta_local[n, ndim]) -> tensor[n] on theta's device.  `theta` is either the full
    [S, ndim] tensor (every rank holds it, e.g. emcee's proposal block) or, with
    theta_is_local=True, just this rank's block of an S = n_samples problem.
    Returns the full [S] tensor on every rank.

    With `ctx` (the white-noise `_lib.Context` that eval_fn would call) on CUDA + NCCL ranks of one box the gather is
    fused into the log-probability kernel (`PeerGather`): no collective is launched; copy=False then returns the
    mapped vector itself, valid until the call after next.  fused=False (or RVLP_PEER_GATHER=0, no `ctx`, a GP
    context, CUDA IPC refused) keeps to ONE NCCL all-gather.  Same bits either way; on 8 B200s the fused step is
    0.03 ms shorter.
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
    if fused is None:
        fused = os.environ.get("RVLP_PEER_GATHER", "1") != "0"
    if (fused and ctx is not None and world > 1 and world <= 8 and getattr(theta, "is_cuda", False)
            and not ctx.desc.is_gp and dist.get_backend(group) == "nccl"):
        dev = theta.device.index if theta.device.index is not None else torch.cuda.current_device()
        pg = _peer_gather_for(ctx, S, dev, group)
        if pg is not None:
            out = pg.logprob(ctx, local, lo)
            return out.clone() if copy else out
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
