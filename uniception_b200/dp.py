"""Data parallelism for the two-view path: one process per GPU, batch of image pairs sharded across
ranks, weights replicated, ONE exchange step -- the all-reduce (average) of the flat fp32 gradient
buffer over NCCL / NVLink.  The reference has no distributed code at all (its users
wrap the modules in DDP, prediction_heads/dpt.py:82); this is the B200-native replacement for that
wrapper on this path.

The flat gradient buffer (params.ParamPack) is reduced in a few large contiguous buckets.  A bucket is
launched on a side stream as soon as the backward pass has finished the parameter range it covers
(heads + decoder while the encoder backward is still running; encoder block groups as they retire),
so the transfer overlaps the remaining backward kernels.  No per-parameter hooks, no Python-side
bucketing copies: the kernels already accumulated into the buffer that is sent.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import torch
import torch.distributed as dist


def shard_batch(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [lo, hi) slice of a batch for `rank`; sizes differ by at most one.  Symmetrized
    partners (a,b),(b,a) sit at 2i, 2i+1 -> shard in units of two to keep them on one rank."""
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_pairs_symmetrized(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    assert n_items % 2 == 0
    lo, hi = shard_batch(n_items // 2, rank, world)
    return 2 * lo, 2 * hi


class GradSync:
    """Overlapped all-reduce(avg) of a flat gradient buffer in prefix-addressed ranges."""

    def __init__(self, flat_grad: torch.Tensor, index: Dict[str, Tuple[int, torch.Size]], group=None,
                 max_bucket_elems: int = 96 * 1024 * 1024, compress_bf16: bool = False):
        """compress_bf16: send every bucket as bf16 (one cast kernel before, one widening copy after the all-reduce): half
        the NVLink bytes and half the HBM traffic the collective takes from the backward kernels, at bf16 precision of the
        AVERAGED gradient (the same trade as DDP's bf16_compress_hook).  Off by default: fp32 matches what the reference's
        users get from DistributedDataParallel."""
        self.flat = flat_grad
        self.compress = bool(compress_bf16) and flat_grad.is_cuda
        self._bf16 = torch.empty(flat_grad.numel(), dtype=torch.bfloat16, device=flat_grad.device) if self.compress else None
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.names = list(index.keys())
        self.offsets = {n: (o, o + s.numel()) for n, (o, s) in index.items()}
        self.max_bucket = max_bucket_elems
        self.side = torch.cuda.Stream() if flat_grad.is_cuda else None
        self.pending: List = []
        self.done_ranges: List[Tuple[int, int]] = []
        self.armed = True  # False inside no_sync(): gradient accumulation micro-steps that must not exchange anything
        self._index = index

    def rebuilt_for(self, flat_grad: torch.Tensor, index: Dict[str, Tuple[int, torch.Size]]) -> "GradSync":
        """A GradSync with the same group / bucket size on a new flat buffer (the ParamPack was rebuilt)."""
        g = GradSync(flat_grad, index, group=self.group, max_bucket_elems=self.max_bucket, compress_bf16=self.compress)
        g.armed = self.armed
        return g

    def no_sync(self):
        """Context manager for gradient accumulation: backward passes inside it only accumulate into the flat buffer
        (`ready()` and `finish()` do nothing), so no all-reduce of micro-step k can race with the kernels of micro-step
        k + 1 that add into the same range.  Run the LAST micro-step outside the context and call `finish()` after it."""
        import contextlib

        @contextlib.contextmanager
        def ctx():
            prev, self.armed = self.armed, False
            try:
                yield self
            finally:
                self.armed = prev

        return ctx()

    def range_of(self, prefix: str) -> Optional[Tuple[int, int]]:
        los = [self.offsets[n] for n in self.names if n.startswith(prefix)]
        if not los:
            return None
        lo, hi = min(a for a, _ in los), max(b for _, b in los)
        return lo, hi

    def _reduce(self, lo: int, hi: int) -> None:
        if self.world == 1 or hi <= lo:
            return
        view = self.flat[lo:hi]
        op = dist.ReduceOp.AVG if self.flat.is_cuda else dist.ReduceOp.SUM
        if self.side is not None:
            self.side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(self.side):
                if self.compress:
                    from . import ops

                    half = self._bf16[lo:hi]
                    ops.cast_bf16(view, out=half)
                    dist.all_reduce(half, op=op, group=self.group)  # on the side stream; ordered before the widening copy
                    view.copy_(half)
                else:
                    self.pending.append(dist.all_reduce(view, op=op, group=self.group, async_op=True))
        else:
            dist.all_reduce(view, op=op, group=self.group)
            view.div_(self.world)

    def ready(self, prefix: str) -> None:
        """The backward pass finished every parameter whose name starts with `prefix`."""
        if not self.armed:
            return
        r = self.range_of(prefix)
        if r is None:
            return
        lo, hi = r
        self.done_ranges.append((lo, hi))
        for s in range(lo, hi, self.max_bucket):
            self._reduce(s, min(hi, s + self.max_bucket))

    def finish(self) -> None:
        """Reduce whatever was not announced, then make the compute stream wait for all buckets."""
        if not self.armed:
            return
        if self.world > 1:
            covered = sorted(self.done_ranges)
            pos = 0
            for lo, hi in covered + [(self.flat.numel(), self.flat.numel())]:
                if lo > pos:
                    for s in range(pos, lo, self.max_bucket):
                        self._reduce(s, min(lo, s + self.max_bucket))
                pos = max(pos, hi)
        for w in self.pending:
            w.wait()
        self.pending.clear()
        self.done_ranges.clear()
        if self.side is not None:
            torch.cuda.current_stream().wait_stream(self.side)


def init_from_env(backend: Optional[str] = None) -> Tuple[int, int, int]:
    """(rank, local_rank, world) from torchrun's environment; initialises the process group when world > 1."""
    import os

    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, local, world
