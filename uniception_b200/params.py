"""Flat parameter storage for a module tree: one fp32 master buffer, one fp32 gradient buffer and one
bf16 shadow buffer (the tensor-core operand copy), with every `nn.Parameter` re-pointed to a view.

Why: (1) the bf16 operand copies of all weights are refreshed by ONE cast kernel per step instead of
one autocast cast per Linear per forward; (2) wgrad kernels accumulate straight into the flat
gradient buffer, which is what the data-parallel all-reduce (dp.py) sends over NVLink in a few large
buckets; (3) weights that the engine wants adjacent (cross-attention projk|projv -> one GEMM) ARE
adjacent, because views follow registration order, which equals the reference's state-dict order.
"""
from __future__ import annotations

from typing import Dict, Tuple

import torch
import torch.nn as nn

from . import ops

_ALIGN = 64  # elements; keeps every view 128-byte aligned in bf16 and 256-byte aligned in fp32


class ParamPack:
    def __init__(self, module: nn.Module):
        named = [(n, p) for n, p in module.named_parameters()]  # de-duplicated by torch
        if not named:
            raise RuntimeError("ParamPack: module has no parameters")
        # Layout order: registration order, except that inside one parent module (e.g. `cross_attn`) all
        # matrices come before all vectors, so projq|projk|projv|proj weights (and their biases) are adjacent.
        ordered, i = [], 0
        while i < len(named):
            key = named[i][0].rsplit(".", 2)[0] if named[i][0].count(".") >= 2 else ""
            j = i
            while j < len(named) and (named[j][0].rsplit(".", 2)[0] if named[j][0].count(".") >= 2 else "") == key:
                j += 1
            grp = named[i:j]
            ordered += [x for x in grp if x[1].dim() >= 2] + [x for x in grp if x[1].dim() < 2]
            i = j
        named = ordered
        dev = named[0][1].device
        if dev.type != "cuda":
            raise RuntimeError("uniception_b200 modules run on CUDA only (no CPU fallback); call .cuda() first")
        self.index: Dict[str, Tuple[int, torch.Size]] = {}
        off = 0
        for n, p in named:
            if p.dtype != torch.float32 or p.device != dev:
                raise RuntimeError(f"ParamPack: parameter {n} must be fp32 on {dev}")
            self.index[n] = (off, p.shape)
            off += (p.numel() + _ALIGN - 1) // _ALIGN * _ALIGN
        self.total = off
        self.flat = torch.zeros(off, dtype=torch.float32, device=dev)
        self.flat_grad = torch.zeros(off, dtype=torch.float32, device=dev)
        self.flat_bf16 = torch.empty(off, dtype=torch.bfloat16, device=dev)
        self.params: Dict[str, nn.Parameter] = {}
        with torch.no_grad():
            for n, p in named:
                o, shp = self.index[n]
                view = self.flat[o:o + p.numel()].view(shp)
                view.copy_(p.data)
                p.data = view
                p.grad = self.flat_grad[o:o + p.numel()].view(shp)
                self.params[n] = p
        self.grad_sync = None  # optional dp.GradSync: told which parameter ranges are final during backward
        self._sentinels = [named[0][1], named[len(named) // 2][1], named[-1][1]]
        self._sentinel_ptrs = [p.data_ptr() for p in self._sentinels]
        self.refresh_bf16()

    # -- validity: .to()/.cuda()/load with assign=True re-create storages and orphan the views
    def valid(self) -> bool:
        return all(p.data_ptr() == q for p, q in zip(self._sentinels, self._sentinel_ptrs))

    def refresh_bf16(self) -> None:
        """One kernel: fp32 masters -> bf16 operand copies (what autocast does per Linear per call)."""
        ops.cast_bf16(self.flat, out=self.flat_bf16)

    def zero_grad(self) -> None:
        self.flat_grad.zero_()

    def w16(self, name: str) -> torch.Tensor:
        o, shp = self.index[name]
        n = 1
        for s in shp:
            n *= s
        return self.flat_bf16[o:o + n].view(shp[0], -1) if len(shp) >= 2 else self.flat_bf16[o:o + n]

    def w16_rows(self, first: str, last: str) -> torch.Tensor:
        """bf16 view spanning two adjacent [r, C] matrices (e.g. projk.weight | projv.weight)."""
        o0, s0 = self.index[first]
        o1, s1 = self.index[last]
        assert s0[1:] == s1[1:] and o1 == o0 + s0.numel(), f"{first} and {last} are not adjacent in the pack"
        return self.flat_bf16[o0:o1 + s1.numel()].view(s0[0] + s1[0], -1)

    def w32(self, name: str) -> torch.Tensor:
        return self.params[name].data

    def w32_span(self, first: str, last: str) -> torch.Tensor:
        o0, s0 = self.index[first]
        o1, s1 = self.index[last]
        assert o1 == o0 + s0.numel(), f"{first} and {last} are not adjacent in the pack"
        return self.flat[o0:o1 + s1.numel()]

    def grad(self, name: str) -> torch.Tensor:
        """fp32 gradient view the kernels accumulate into (2-D for matrices)."""
        o, shp = self.index[name]
        g = self.flat_grad[o:o + shp.numel()]
        return g.view(shp[0], -1) if len(shp) >= 2 else g

    def grad_span(self, first: str, last: str) -> torch.Tensor:
        o0, s0 = self.index[first]
        o1, s1 = self.index[last]
        return self.flat_grad[o0:o1 + s1.numel()]

    def requires_grad(self, name: str) -> bool:
        return self.params[name].requires_grad

    def notify_done(self, prefix: str) -> None:
        if self.grad_sync is not None:
            self.grad_sync.ready(prefix)

    def rebind_grads(self) -> None:
        """Re-attach `.grad` views (e.g. after `zero_grad(set_to_none=True)` dropped them)."""
        for n, p in self.params.items():
            if p.grad is None or p.grad.data_ptr() != self.grad(n).data_ptr():
                o, shp = self.index[n]
                p.grad = self.flat_grad[o:o + shp.numel()].view(shp)


def get_pack(module: nn.Module) -> ParamPack:
    """The pack of `module`, (re)built lazily when parameters moved."""
    pk = module.__dict__.get("_uc_pack")
    if pk is None or not pk.valid():
        pk = ParamPack(module)
        module.__dict__["_uc_pack"] = pk
    return pk
