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
                p._uc_pack = (self, n)  # lets code that assigns gradients itself (dpt_engine.ConvW.flush_grads) find the flat view
                self.params[n] = p
        self.grad_sync = None  # optional dp.GradSync: told which parameter ranges are final during backward
        self._sentinels = [named[0][1], named[len(named) // 2][1], named[-1][1]]
        self._sentinel_named = [named[0], named[len(named) // 2], named[-1]]
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

    def grad_view(self, name: str) -> torch.Tensor:
        """The flat-buffer gradient view of one parameter, in the parameter's own shape."""
        o, shp = self.index[name]
        return self.flat_grad[o:o + shp.numel()].view(shp)

    def prepare_grads(self) -> None:
        """Called at the start of every backward node.  `optimizer.zero_grad()` / `Module.zero_grad()` default to
        set_to_none=True: a parameter whose `.grad` was dropped restarts from zero -- ONLY its own range of the flat
        buffer is cleared (one memset when every gradient was dropped), so parameters that are frozen or simply not in
        the optimizer keep their state.  A `.grad` tensor that somebody else assigned (not a view of the flat buffer) is
        folded into the view instead of being discarded.  The per-call cost with nothing to do is one attribute read per
        parameter."""
        dropped = [n for n, p in self.params.items() if p.grad is None]
        if dropped:
            if len(dropped) == len(self.params):
                self.flat_grad.zero_()
            else:
                for n in dropped:
                    self.grad_view(n).zero_()
            for n in dropped:
                self.params[n].grad = self.grad_view(n)
        if dropped or any(p.grad is not None and p.grad.data_ptr() != self.flat_grad.data_ptr() + 4 * self.index[n][0]
                          for n, p in self._sentinel_named):
            self.rebind_grads()

    def rebind_grads(self) -> None:
        """Re-attach `.grad` views; a foreign `.grad` (assigned by other code) is accumulated into the view first."""
        base = self.flat_grad.data_ptr()
        for n, p in self.params.items():
            want = base + 4 * self.index[n][0]
            if p.grad is None:
                p.grad = self.grad_view(n)
            elif p.grad.data_ptr() != want:
                view = self.grad_view(n)
                view.add_(p.grad.to(view.dtype).view(view.shape))
                p.grad = view


def get_pack(module: nn.Module) -> ParamPack:
    """The pack of `module`, (re)built lazily when parameters moved."""
    pk = module.__dict__.get("_uc_pack")
    if pk is None or not pk.valid():
        old = pk
        pk = ParamPack(module)
        if old is not None and old.grad_sync is not None:
            # the parameters moved (.to() / .cuda() / load with assign=True): data parallelism must keep synchronising, on the
            # NEW flat gradient buffer
            pk.grad_sync = old.grad_sync.rebuilt_for(pk.flat_grad, pk.index)
        module.__dict__["_uc_pack"] = pk
    return pk
